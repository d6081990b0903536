// Explicit instantiation of every kernel launcher for ONE polynomial size (compile with -DBMI_INST_L=10..14).
#include "launchers.cuh"

#ifndef BMI_INST_L
#error "compile with -DBMI_INST_L=<log2 of the polynomial size>"
#endif
template int setup_attrs<BMI_INST_L>(const bmi_ctx*);
template int launch_convert<BMI_INST_L>(bmi_ctx*, const u64*, u64* const*, int64_t, int64_t, cudaStream_t);
template int launch_pbs<BMI_INST_L>(bmi_ctx*, PbsArgs, cudaStream_t);
template int launch_polymul<BMI_INST_L>(bmi_ctx*, const u64*, const u64*, u64*, int, cudaStream_t);
template int64_t split_capacity<BMI_INST_L>(bmi_ctx*);
