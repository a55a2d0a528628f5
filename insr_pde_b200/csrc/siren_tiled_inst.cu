// the tiled family in its own object (built in parallel with the other units)
#include "siren_tiled.cuh"
#include "siren_tiled_inst.cuh"
