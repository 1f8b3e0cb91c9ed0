"""nafp-b200: B200-native inference fingerprinting + retrieval path of neural-audio-fp.

Host side (Python) mirrors the reference's call surfaces:
  model.generate.generate_fingerprint / test_step     <- reference model/generate.py
  eval.utils.get_index.get_index (Index objects)       <- reference eval/utils/get_index_faiss.py
  eval.eval_search.eval_faiss (click CLI)              <- reference eval/eval_faiss.py
and calls hand-written sm_100a kernels through the C ABI declared in include/nafp.h
(csrc/libnafp.so, loaded with ctypes by ``_lib``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
