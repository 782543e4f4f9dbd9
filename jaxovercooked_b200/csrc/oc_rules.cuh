// oc_rules.cuh: small device helpers and process_interact for one faced cell -- part of the sm_100a Overcooked kernels (see oc_b200.cu for the overview)
#pragma once
#include <cstdint>

namespace {

// ------------------------------------------------------------------------------------------ small helpers

__device__ __forceinline__ unsigned fastdiv(unsigned x, unsigned magic) { return __umulhi(x, magic); }
__device__ __forceinline__ int dir_x(int d) { return (d == 2) - (d == 3); }
__device__ __forceinline__ int dir_y(int d) { return (d == 1) - (d == 0); }
__device__ __forceinline__ int obj_colour(int o) { return (int)((0x06716644500ull >> (4 * o)) & 0xFull); }  // common.py:51-63
__device__ __forceinline__ int inv_code(int inv) { return inv == 9 ? 3 : ((inv >> 1) & 3); }               // 1,3,5,9 -> 0..3
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Interact { int obj, status, inv, delivery, shaped, write; };

// overcooked.py:516-642 for one agent facing a cell holding (obj, status); `wall` = wall_map at that cell.
__device__ __forceinline__ Interact interact_cell(int obj, int status, int inv, bool wall) {
    const bool is_pot = obj == 8, is_goal = obj == 7, is_agent = obj == 10;
    const bool is_pile = (obj == 6) | (obj == 4);
    const bool pickable = (obj == 5) | (obj == 3) | (obj == 9);
    const bool is_table = wall & !is_pot;
    const bool table_empty = (obj == 2) | (obj == 1);
    const bool case1 = is_pot & (status > 20) & (inv == 3);      // onion into a pot that is not full
    const bool case2 = is_pot & (status == 0) & (inv == 5);      // plate a finished soup
    const bool pickup = is_table & !table_empty & (inv == 1) & (is_pile | pickable);
    const bool drop = is_table & table_empty & (inv != 1);
    const bool delivery = is_table & is_goal & (inv == 9);
    Interact r;
    r.status = case1 ? status - 1 : (case2 ? 23 : status);
    int ninv = case1 ? 1 : (case2 ? 9 : inv);
    r.obj = obj;
    if (pickup) { r.obj = is_pile ? obj : 2; ninv = (obj == 6) ? 5 : ((obj == 4) ? 3 : obj); }
    if (drop) { r.obj = inv; ninv = 1; }
    if (delivery) ninv = 1;
    r.inv = ninv;
    r.delivery = delivery;
    // :568-569; the plate-pickup bonus (:616-628) is identically zero: the ring's colour byte is 5
    r.shaped = case1 ? 3 : (case2 ? 5 : 0);
    r.write = !is_agent;
    return r;
}

}  // namespace
