// [TF1] tf.train.AdamOptimizer (models/DAEs.py:102, :198) == ApplyAdam functor, shared by the dense
// optimizer kernels (optim.cu) and the fused dW + Adam epilogue (gemm_sm100.cu):
//     m   += (g - m) * (1 - beta1)
//     v   += (g*g - v) * (1 - beta2)
//     var -= (m * alpha) / (sqrt(v) + eps),   alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
// Every operation is an explicitly rounded fp32 op (no FMA contraction) so the update is bit-exact
// against the NumPy oracle given the same gradient.
#pragma once
#include <cuda_runtime.h>

namespace dae {

struct AdamConst {
    float alpha, omb1, omb2, eps, lambda;
};

__device__ __forceinline__ void adam_one(float& w, float& m, float& v, float g, const AdamConst c) {
    if (c.lambda != 0.f) g = __fadd_rn(g, __fmul_rn(c.lambda, w));          // d/dw of lambda * l2_loss(w)  (DAEs.py:100)
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), c.omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), c.omb2));
    // m == 0 (an element that has never seen a gradient: catalogue rows no playlist has listed yet, dead CNN features,
    // padding columns): the step is (+0 * alpha) / (sqrt(v) + eps) = +0 and w - (+0) == w bit for bit -- but the
    // correctly rounded sqrt and division take their SLOW paths for a zero operand (measured: +56 % instructions in the
    // title output layer's update, which made that HBM-bound kernel compute-bound).  m is +0, never -0, in that case:
    // it starts as +0 and +0 + ((+-0) - (+0)) * (1 - beta1) == +0.
    if (m != 0.f) w = __fsub_rn(w, __fdiv_rn(__fmul_rn(m, c.alpha), __fadd_rn(__fsqrt_rn(v), c.eps)));
}

}  // namespace dae
