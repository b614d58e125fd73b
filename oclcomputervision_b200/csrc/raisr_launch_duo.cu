// raisr_launch_duo.cu -- placeholder translation unit of the two-types-per-CTA filter kernel (raisr_duo.cuh).
#include "raisr_internal.h"
