"""visuelle2-multimodal-fusion_b200: B200-native hot path of the Visuelle 2.0 multimodal forecasters.

Drop-in ``nn.Module``s with the reference's constructors / forward signatures / state_dict keys
(``models/``) whose bodies call hand-written sm_100a CUDA through the C-ABI in ``include/v2f.h``
(``csrc/`` -> ``libv2f_b200.so``, loaded with ctypes by ``_lib.py``).  There is no CPU path:
importing the package is cheap, but any compute call raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
