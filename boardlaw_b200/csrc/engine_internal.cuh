// Internal interfaces between the engine translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

// bl_tree.counters slots
enum { C_EVALS = 0, C_CHILDREN = 1, C_ITERS = 2, C_DESCENTS = 3, C_BACKUP_NODES = 4, C_ERRORS = 5, C_MOVE = 6, C_QUEUE = 7 };

// descend.cu: task-parallel descent (writes t.leaf = existing terminal child or -1, t.leaf_parent, t.leaf_action)
// followed by expand + env step.  Returns a cudaError_t / negative argument error.
int bl_descend_v2(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
