// rmx_k_fwd_i2_w2_g2a.cu -- explicit instances of rollout_fwd_kernel (see rmx_host.h): IMPL, NW, GROUND, ADJ, LIN
#include "rmx_launch.cuh"
#define X(IMPL, NW, G, A, L) RMX_DEFINE_FWD(IMPL, NW, G, A, L)
X(2,2,2,1,0)
