// placeholder (tcgen05 path is added next)
#include "common.cuh"
