/* transport_strict.cu -- parity flavour of the transport kernels: compile with -fmad=false. */
#define ALVRL_FLAVOR strict
#include "transport.cuh"
#include "kernels.h"
#include "transport_launch.inl"
