// Stubs for the measured-and-rejected descent variants (descend_pc.cu = 4, descend_all.cu = 6, descend_pk.cu = 7) when the library is built without
// BL_EXPERIMENTAL=1 (boardlaw_b200/build.py): the dispatcher sees "unsupported" and bl_debug_set_descend_variant refuses them.
#include "engine_internal.cuh"

int bl_descend_pc(const bl_tree *, int, const bl_half *, uint64_t, cudaStream_t) { return -2; }
int bl_descend_pk(const bl_tree *, int, const bl_half *, uint64_t, cudaStream_t) { return -2; }
int bl_descend_all(const bl_tree *, int, const bl_half *, uint64_t, cudaStream_t) { return -2; }
int64_t bl_all_scratch_bytes(const bl_tree *) { return 0; }
bool bl_experimental_built() { return false; }
