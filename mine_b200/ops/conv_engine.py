"""sm_100a convolution engine (tcgen05/TMEM implicit GEMM) - see csrc/conv_tcgen05.cu."""
AVAILABLE = False
