// Host-side entry points of the C ABI that launch nothing: version, device selection, exp table.
#include <math.h>
#include <string.h>

#include "common.cuh"

extern "C" int bl_abi_version(void) { return 9; }

// The library links the CUDA runtime statically, so it carries its own "current device"; the host
// wrapper calls this with the device index of the tensors before launching (the reference does the
// same with a CUDAGuard, boardlaw/hex/cpp/cuda.cu:140).
extern "C" int bl_set_device(int device) { return (int)cudaSetDevice(device); }

// expf over every binary16 bit pattern, evaluated by the host libm: the reference's CPU build takes
// exp(logit) from libm (boardlaw/mcts/cpp/cpu.cpp:89), and a table is the only way to be bit-identical
// to it on every host (glibc picks FMA / non-FMA variants per CPU).
extern "C" void bl_exp_table_host(float *out) {
    for (uint32_t i = 0; i < 65536; i++) {
        uint16_t h = (uint16_t)i;
        _Float16 x;
        memcpy(&x, &h, 2);
        volatile float f = (float)x;
        out[i] = expf(f);
    }
}
