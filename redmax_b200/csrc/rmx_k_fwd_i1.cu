// rmx_k_fwd_i1.cu -- explicit instances of rollout_fwd_kernel (see rmx_host.h): IMPL, NW, GROUND, ADJ, LIN
#include "rmx_launch.cuh"
#define X(IMPL, NW, G, A, L) RMX_DEFINE_FWD(IMPL, NW, G, A, L)
X(1,1,0,0,0) X(1,1,1,0,0) X(1,1,0,1,0) X(1,1,1,1,0) X(1,2,0,0,0) X(1,2,1,0,0) X(1,2,0,1,0) X(1,2,1,1,0) X(1,4,0,0,0) X(1,4,1,0,0) X(1,4,0,1,0) X(1,4,1,1,0)
