"""lnasr -- B200-native drop-in for the acoustic hot path of yehuohan/ln-asr.

Same modules and classes as the reference package (`lnasr.mfcc.MFCC`, `lnasr.gmm`, `lnasr.hmm.HMM`,
`lnasr.gmmhmm.GMMHMM`, `lnasr.utils`), computed by hand-written sm_100a CUDA kernels behind the C ABI of
`liblnasr_b200.so` (include/lnasr_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
