// rmx_adjoint.cuh -- adjoint forward (tape) and backward kernels (filled in below)
#pragma once
#include "rmx_rollout.cuh"
