/* transport_fast.cu -- speed flavour of the transport kernels (FMA contraction + MUFU intrinsics). */
#define ALVRL_FLAVOR fast
#define ALVRL_FAST 1
#include "transport.cuh"
#include "kernels.h"
#include "transport_launch.inl"
