"""Artefact IO (mirror of mobilellm/utils/io.py:34-69): JSON written with indent=4, sort_keys=True."""
import json, logging, os, sys, time


def json_load(path):
    with open(path, "r") as f:
        return json.load(f)


def json_save(path, obj):
    with open(path, "w") as f:
        json.dump(obj, f, indent=4, sort_keys=True)


def create_logger(output_dir=None, dist_rank=0, name=""):
    logger = logging.getLogger(name or "mobilequant_b200")
    logger.setLevel(logging.INFO)
    logger.propagate = False
    if logger.handlers:
        return logger
    fmt = "[%(asctime)s %(name)s] (%(filename)s %(lineno)d): %(levelname)s %(message)s"
    if dist_rank == 0:
        h = logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter(fmt=fmt, datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(h)
    if output_dir is not None:
        os.makedirs(str(output_dir), exist_ok=True)
        fh = logging.FileHandler(os.path.join(str(output_dir), f"log_rank{dist_rank}_{int(time.time())}.txt"), mode="a")
        fh.setFormatter(logging.Formatter(fmt=fmt, datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(fh)
    return logger
