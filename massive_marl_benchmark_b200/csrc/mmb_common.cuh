// mmb_common.cuh - host-side helpers shared by the translation units of libmmb_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MMB_TEN_ANT_EPT
#define MMB_TEN_ANT_EPT 32  // environments per CTA tile of the TenAnt kernel (multiple of 4)
#endif
#define MMB_MAX_DEVICES 16

namespace mmb {
void count_launch();  // api.cu: bumps the process-wide launch counter reported by mmb_launch_count()
}
