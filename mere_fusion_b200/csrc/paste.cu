#include "mf_common.cuh"
