// anoddpm_b200/csrc/diffusion_step.cuh -- the posterior step shared by p_sample_kernel (diffusion.cu) and the fused
// noise + posterior kernel (simplex.cu).  Both translation units are compiled with -fmad=false: the reference's operation order,
// no fma contraction (GaussianDiffusion.py:228-230, 253-267, 282-318).
#pragma once

__device__ __forceinline__ float p_step(float xt, float e, float z, float ra, float rm1, float c1, float c2, float sg,
                                        float& x0_out) {
    float x0 = ra * xt - rm1 * e;
    x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    x0_out = x0;
    const float mean = c1 * x0 + c2 * xt;
    return mean + sg * z;
}
