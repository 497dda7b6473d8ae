// decimate_inst.cu -- the instantiations of mix_decimate_fast_kernel for ONE type pair (-DDF_IN=0|1|2 -DDF_OUT=0|1; 2 = already mixed): six
// translation units built in parallel (the 32 instantiations in one unit took three minutes of nvcc).
#include "decimate_kernels.cuh"

#ifndef DF_IN
#error "compile with -DDF_IN=0|1 -DDF_OUT=0|1"
#endif

#define DF_CAT2(a, b, c) a##b##_##c
#define DF_CAT(a, b, c) DF_CAT2(a, b, c)

namespace dmix {

// kernel for (CTA of 256 / 128 threads, filter shape) of this unit's type pair
DecimFastKernel DF_CAT(df_kernel_, DF_IN, DF_OUT)(int nt128, int shape)
{
    static const DecimFastKernel k[2][4] = {
        {mix_decimate_fast_kernel<DF_IN, DF_OUT, 0, 256>, mix_decimate_fast_kernel<DF_IN, DF_OUT, 1, 256>,
         mix_decimate_fast_kernel<DF_IN, DF_OUT, 2, 256>, mix_decimate_fast_kernel<DF_IN, DF_OUT, 3, 256>},
        {mix_decimate_fast_kernel<DF_IN, DF_OUT, 0, 128>, mix_decimate_fast_kernel<DF_IN, DF_OUT, 1, 128>,
         mix_decimate_fast_kernel<DF_IN, DF_OUT, 2, 128>, mix_decimate_fast_kernel<DF_IN, DF_OUT, 3, 128>}};
    return k[nt128 ? 1 : 0][shape & 3];
}

}  // namespace dmix
