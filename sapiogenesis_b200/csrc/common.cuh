// common.cuh -- shared device helpers for the sm_100a kernels of the Sapiogenesis world step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sapio.h"

#define MAXC SAPIO_MAXC
#define FULL 0xffffffffu

// The world travels to every kernel by value (pointers + frozen params, ~1.2 KB of kernel parameters).
#define WORLD_ARG const __grid_constant__ SapioWorld w

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- exact geometric predicates -------------------------------------------------------------------------------------
// Contact / view membership must be bit-exact against the oracle, which evaluates them in double on the fp32
// positions (SURVEY A-3). fp32 products are exact in double; explicit _rn intrinsics forbid FMA contraction so the
// operation sequence is the oracle's.
__device__ __forceinline__ bool circles_overlap(float x1, float y1, float r1, float x2, float y2, float r2, double *dsq) {
    double dx = __dsub_rn((double)x2, (double)x1), dy = __dsub_rn((double)y2, (double)y1);
    double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    double md = __dadd_rn((double)r1, (double)r2);
    *dsq = d2;
    return d2 < __dmul_rn(md, md);
}

struct WallHit { float nx, ny, cx, cy; double d; };   // normal (circle -> wall), closest point on the segment, centre distance

// cpCollision.c CircleToSegment against wall wi (physics.py:328-338: 0 left, 1 bottom, 2 right, 3 top), same double
// operation sequence as oracle/space_step.c:circle_wall.
__device__ __forceinline__ bool circle_wall(const SapioParams &p, int wi, float cxf, float cyf, float rf, WallHit *h) {
    double fw = (double)p.frame_width, W = (double)p.width, H = (double)p.height;
    double ax, ay, bx, by;
    double x0 = fw, y0 = fw, x1 = __dsub_rn(W, fw), y1 = __dsub_rn(H, fw);
    switch (wi) {
        case 0: ax = x0; ay = y0; bx = x0; by = y1; break;
        case 1: ax = x0; ay = y1; bx = x1; by = y1; break;
        case 2: ax = x1; ay = y0; bx = x1; by = y1; break;
        default: ax = x0; ay = y0; bx = x1; by = y0; break;
    }
    double sdx = __dsub_rn(bx, ax), sdy = __dsub_rn(by, ay);
    double cx = (double)cxf, cy = (double)cyf;
    double num = __dadd_rn(__dmul_rn(sdx, __dsub_rn(cx, ax)), __dmul_rn(sdy, __dsub_rn(cy, ay)));
    double den = __dadd_rn(__dmul_rn(sdx, sdx), __dmul_rn(sdy, sdy));
    double t = __ddiv_rn(num, den);
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    double qx = __dadd_rn(ax, __dmul_rn(sdx, t)), qy = __dadd_rn(ay, __dmul_rn(sdy, t));
    double mind = __dadd_rn((double)rf, fw);
    double ddx = __dsub_rn(qx, cx), ddy = __dsub_rn(qy, cy);
    double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
    if (!(d2 < __dmul_rn(mind, mind))) return false;
    double d = __dsqrt_rn(d2);
    double nx, ny;
    if (d != 0.0) { double inv = __ddiv_rn(1.0, d); nx = __dmul_rn(ddx, inv); ny = __dmul_rn(ddy, inv); }
    else { double L = __dsqrt_rn(den); double inv = __ddiv_rn(1.0, L); nx = -__dmul_rn(sdy, inv); ny = __dmul_rn(sdx, inv); }
    h->nx = (float)nx; h->ny = (float)ny; h->cx = (float)qx; h->cy = (float)qy; h->d = d;
    return true;
}

__device__ __forceinline__ bool is_sensor_type(int t) { return t == SAPIO_T_EYE || t == SAPIO_T_OLFACTORY; }

// index of pin joint (i,j), i<j, in lexicographic order == Chipmunk constraint insertion order (sprites.py:1721-1726)
__host__ __device__ __forceinline__ int pair_index(int i, int j) { return i * (2 * MAXC - i - 1) / 2 + (j - i - 1); }
