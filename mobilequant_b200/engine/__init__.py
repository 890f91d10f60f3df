from .int_forward import IntEngine
