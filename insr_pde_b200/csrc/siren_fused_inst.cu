// one object per (D, O): nvcc -DINSR_INST_D=<D> -DINSR_INST_O=<O> -c siren_fused_inst.cu
#include "siren_fused.cuh"
#include "siren_tc.cuh"
#include "siren_fused_inst.cuh"
