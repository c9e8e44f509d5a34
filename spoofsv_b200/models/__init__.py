from .TTSModel import SSRN, highwayConv, melSyn  # noqa: F401
