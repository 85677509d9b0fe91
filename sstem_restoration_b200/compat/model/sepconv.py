# Drop-in for sff_scripts_interp/model/sepconv.py (the CuPy variant) of ssTEM-restoration.
from sstem_restoration_b200.sepconv import FunctionSepconv, ModuleSepconv, _FunctionSepconv  # noqa: F401
