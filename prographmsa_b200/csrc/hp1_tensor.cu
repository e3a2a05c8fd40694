// hp1_tensor.cu -- K4 (tensor flavour): placeholder until the tcgen05 kernel lands in this file.
#include "internal.h"
int launch_emission_tensor(cudaStream_t, int, int, const GraphJobDev*, const GraphJobDev*) {
    pgm_set_error("PGM_EMISSION_TENSOR is not built yet");
    return PGM_ERR_ARG;
}
