"""mobilequant_b200 -- B200-native implementation of MobileQuant's two hot paths.

Layout mirrors the reference's package so the parity tests read like the reference's own call sites:
  mobilequant_b200.quantization.qmodule    <-> mobilellm/quantization/qmodule.py
  mobilequant_b200.quantization.algorithm  <-> mobilellm/quantization/algorithm.py
  mobilequant_b200.model.hf_model          <-> mobilellm/model/hf_model.py
  mobilequant_b200.engine                  --  the statically-quantised integer forward (new; the reference hands this
                                               to Qualcomm QNN)
"""
__version__ = "0.1.0"
