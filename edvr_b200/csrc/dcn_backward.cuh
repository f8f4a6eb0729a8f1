// dcn_backward.cuh — DCNv2 backward (placeholder until the first forward GPU check passes).
#pragma once
#include "common.cuh"

namespace eb {

struct DcnBwdParams {
    const float *x, *offset, *mask, *weight, *grad_out;
    float *grad_x, *grad_offset, *grad_mask, *grad_weight, *grad_bias, *gcol;
    int N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, Ho, Wo;
};

inline int dcn_backward_launch(const DcnBwdParams&, cudaStream_t, int) { return EB_ERR_UNSUPPORTED; }

}  // namespace eb
