"""
kir_graph_b200: B200-native allele-typing core for Graph-KIR.

The package keeps the reference's Python surface for the typing path
(``typing_mulit_allele.AlleleTyping``, ``kir_typing.selectKirTypingModel`` ...)
and routes the arithmetic through a C-ABI CUDA library
(``include/gk_typing.h`` -> ``kir_graph_b200/lib/libgk_typing.so``).
There is no CPU fallback: the typing classes raise if the CUDA library or a
GPU is missing.
"""

__version__ = "0.1.0"
