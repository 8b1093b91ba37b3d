"""B200-native triple-hybrid retrieval scoring path (dense top-k, BM25 top-k, weighted RRF fusion
with safety/denoise filters, MaxSim rerank) behind the reference's retriever call surface.

The compute path is libthr.so (hand-written sm_100a CUDA, C-ABI in include/thr.h); importing
this package does not load it, constructing an Engine does — and fails loudly without a GPU.
"""
__version__ = "0.1.0"
