"""spoofsv_b200 -- B200-native Text2Mel + SSRN synthesis hot path of MingruiYuan/SpoofSV.

Only what the hot path needs: hand-written sm_100a CUDA behind a C ABI
(``csrc/``, ``include/spoofsv_b200.h``), the drop-in ``melSyn`` / ``SSRN`` modules
(``models/TTSModel.py``) and the synthesis drivers.  No CPU fallback exists.
"""
__version__ = "0.1.0"
