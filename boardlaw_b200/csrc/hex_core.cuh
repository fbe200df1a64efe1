// Per-lane Hex rules on a board staged in shared memory.
// Semantics follow the reference's step kernel (boardlaw/hex/cpp/cuda.cu:18-137) — see SURVEY.md A.2.
#pragma once
#include "common.cuh"

enum : uint8_t { BL_EMPTY = 0, BL_BLACK = 1, BL_WHITE = 2, BL_TOP = 3, BL_BOT = 4, BL_LEFT = 5, BL_RIGHT = 6 };

// Neighbour offsets in the reference's order (cuda.cu:22,99); the order only matters for the
// priority of the off-board tests, which is row-before-column below.
#define BL_NBR_DR(k) ((k) < 2 ? -1 : ((k) < 4 ? 0 : 1))
#define BL_NBR_DC(k) ((k) == 0 ? 0 : (k) == 1 ? 1 : (k) == 2 ? -1 : (k) == 3 ? 1 : (k) == 4 ? -1 : 0)

// Places `action` (mover's frame) for `seat` on the lane's board `bd` (cell c at bd[c*pitch]),
// relabels the connected plain group when it touches an edge group, and returns
// 0 = no win, 1 = seat 0 (black) won, 2 = seat 1 (white) won.  `stk` is a scratch column of A bytes
// (or 16-bit cells when A > 255) with the same pitch.
template <typename StkT>
__device__ __forceinline__ int bl_hex_place(uint8_t *bd, StkT *stk, int pitch, int S, int seat, int action) {
    int row, col;
    if (seat == 0) { row = action / S; col = action - row * S; }     // cuda.cu:88-91: white plays transposed
    else           { col = action / S; row = action - col * S; }

    unsigned adj = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int r = row + BL_NBR_DR(k), c = col + BL_NBR_DC(k);
        unsigned code;
        if (r < 0) code = BL_TOP;
        else if (r >= S) code = BL_BOT;
        else if (c < 0) code = BL_LEFT;
        else if (c >= S) code = BL_RIGHT;
        else code = bd[(r * S + c) * pitch];
        adj |= 1u << code;
    }
    int win = 0;
    uint8_t plain, new_val;
    if (seat) {
        plain = BL_WHITE;
        if ((adj >> BL_LEFT & 1) && (adj >> BL_RIGHT & 1)) win = 2;
        new_val = (adj >> BL_LEFT & 1) ? BL_LEFT : ((adj >> BL_RIGHT & 1) ? BL_RIGHT : BL_WHITE);
    } else {
        plain = BL_BLACK;
        if ((adj >> BL_TOP & 1) && (adj >> BL_BOT & 1)) win = 1;
        new_val = (adj >> BL_TOP & 1) ? BL_TOP : ((adj >> BL_BOT & 1) ? BL_BOT : BL_BLACK);
    }
    int start = row * S + col;
    bd[start * pitch] = new_val;       // plain colour if no edge contact, else already the flooded label
    if (new_val >= BL_TOP) {
        // relabel the 6-connected component of `plain` cells around the new stone (flood, cuda.cu:18-74);
        // cells are relabelled when pushed, so every cell enters the stack at most once.
        int sp = 0;
        stk[0] = (StkT)start;
        sp = 1;
        while (sp) {
            int cell = stk[(--sp) * pitch];
            int r0 = cell / S, c0 = cell - r0 * S;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                int r = r0 + BL_NBR_DR(k), c = c0 + BL_NBR_DC(k);
                if (r >= 0 && r < S && c >= 0 && c < S) {
                    int nb = r * S + c;
                    if (bd[nb * pitch] == plain) {
                        bd[nb * pitch] = new_val;
                        stk[sp * pitch] = (StkT)nb;
                        sp++;
                    }
                }
            }
        }
    }
    return win;
}
