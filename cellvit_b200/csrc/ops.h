// ops.h -- host launchers of the non-GEMM kernels (layer norm, attention, layout transforms, stencils).
#pragma once
#include "common.cuh"
