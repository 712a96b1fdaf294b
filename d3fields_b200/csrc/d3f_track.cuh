// d3f_track.cuh — one Adam iteration of rigid_tracking (reference fusion.py:1643-1665) as ONE launch.
//
// The four-launch iteration (d3f_eval -> d3f_track_loss_grad -> d3f_eval_backward -> d3f_track_update) spends most of its
// 26 us on launch boundaries and on reading the same 16 texel rows of a point twice (forward, then backward), with the
// descriptor row and its gradient making a round trip through memory in between.  A tracking step is a few hundred points
// (fusion.py:1650): latency, not bandwidth.  Here one CTA owns one point for the whole iteration:
//
//   pts_p = last_p @ R(log_r) + t                                   (every CTA, from the instance's parameters)
//   per-view projection / visibility / weight                        (thread v, view_sample: as d3f_eval)
//   the 4 views x 4 corners texel rows -> REGISTERS, one float4 per thread and row: all 16 loads in flight at once,
//   feat = sum_v sum_corner (w_corner * weight_v / (count+1e-6)) texel      (d3f_eval's order: views, then nw ne sw se)
//   |feat - src|_2  (block reduction)  ->  d loss / d feat  in registers
//   sum_c g_c d feat_c / d(ix, iy, weight) per view from the SAME registers (block reduction)
//   chain to the camera frame and the point (thread v), d loss / d pts_p, the point's loss term
//   arrive on the instance's counter; the LAST CTA of an instance reduces d loss / d pts over the instance's points and
//   takes the Adam step (track_update_instance): the next launch starts from (t_out, r_out).
//
// Texels are read once, the descriptor row and its gradient never leave the SM, and an iteration is one launch.
// Limits (else the caller uses the four-launch path): V <= 4, float32 map with C % 4 == 0, C <= 1024, 16-byte aligned rows.
#pragma once
#include "d3f_common.cuh"
#include "d3f_aux.cuh"

namespace d3f {

constexpr int STEP_THREADS = 256;
constexpr int STEP_VIEWS = 4;
constexpr int STEP_MAX_C = STEP_THREADS * 4;

struct TrackStepParams {
    const float* __restrict__ pose;      // observation, as D3FObs
    const float* __restrict__ K;
    const float* __restrict__ depth;
    int32_t V, H, W;
    float mu;
    const float* __restrict__ map;       // (V,h,w,C) float32 descriptor map
    int32_t h, w, C;
    int64_t sv;
    int32_t sy, sx;
    const float* __restrict__ src;       // (I*P, C) descriptors to match
    float dist_w;
    float* __restrict__ grad_pts;        // (I*P,3) scratch: d loss / d pts of this launch
    float* __restrict__ loss_terms;      // (I*P) or nullptr
    unsigned* __restrict__ arrivals;     // (I) zero before the first launch; every launch leaves it zero
    TrackParams tp;                      // tp.grad_pts is ignored (the scratch above is used); tp.pts: points of THIS launch (nullable)
};

// what thread v leaves for the CTA about view v
struct StepView {
    float px, py, cz, d, weight;
    float wx, wy;                        // fractional position in the texel cell
    float m[4];                          // 1 where the corner is inside the map (nw ne sw se)
    float fw[4];                         // d3f_eval's corner weights (0 outside)
    int32_t o[4];                        // element offsets of the four (clamped) corner rows inside the view
    int32_t vis;
};

// MINB: resident CTAs per SM the register budget is cut for — 2 (105 registers, no spills; 296 CTAs in flight) or
// 3 (80 registers, twenty spilled words; 444 in flight: the reference's 4 objects x 100 points fit in one wave)
template <bool RECIP, int MINB>
__global__ void __launch_bounds__(STEP_THREADS, MINB)
track_step_kernel(const TrackStepParams sp) {
    __shared__ float sH[STEP_VIEWS * 12];
    __shared__ StepView s_view[STEP_VIEWS];
    __shared__ float sR[9], sT[3];
    __shared__ float s_red[STEP_THREADS / 32][TRACK_RED];
    __shared__ float s_ss[STEP_THREADS / 32];
    __shared__ float s_c[STEP_VIEWS][3];
    __shared__ int s_last;
    const TrackParams& tp = sp.tp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = sp.V;
    const int i = blockIdx.x, inst = i / tp.P;
    const int n = tp.I * tp.P;

    // requested first, used later: the point of the previous frame and this thread's four channels of the descriptor to
    // match — their round trip overlaps the projection set-up
    const float* last = tp.last_pts + (size_t)i * 3;
    const float lx = __ldg(last), ly = __ldg(last + 1), lz = __ldg(last + 2);
    const int c = tid * 4;
    const bool active = c < sp.C;
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) s4 = __ldg(reinterpret_cast<const float4*>(sp.src + (size_t)i * sp.C + c));

    if (tid < V * 3) {
        const int v = tid / 3, k = tid - v * 3;
        float row[4];
        krt_row(sp.K + v * 9, sp.pose + v * 12, k, row);
        sH[v * 12 + k * 4 + 0] = row[0]; sH[v * 12 + k * 4 + 1] = row[1];
        sH[v * 12 + k * 4 + 2] = row[2]; sH[v * 12 + k * 4 + 3] = row[3];
    } else if (tid == 32) {                              // another warp: the instance's rotation and translation
        const float w3[3] = {tp.r_in[inst * 3], tp.r_in[inst * 3 + 1], tp.r_in[inst * 3 + 2]};
        float a2, f1, f2;
        so3_exp(w3, sR, a2, f1, f2);
        sT[0] = tp.t_in[inst * 3]; sT[1] = tp.t_in[inst * 3 + 1]; sT[2] = tp.t_in[inst * 3 + 2];
    }
    __syncthreads();
    // pts = last @ R + t, the expression of track_update_kernel
    const float x = lx * sR[0] + ly * sR[3] + lz * sR[6] + sT[0];
    const float y = lx * sR[1] + ly * sR[4] + lz * sR[7] + sT[1];
    const float z = lx * sR[2] + ly * sR[5] + lz * sR[8] + sT[2];

    if (tid < V) {                                       // per-view forward, one thread per view
        const int v = tid;
        float Hm[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) Hm[j] = sH[v * 12 + j];
        const ViewSample sm = view_sample<RECIP>(Hm, x, y, z, sp.depth + (size_t)v * sp.H * sp.W, sp.H, sp.W, sp.mu, false);
        StepView& q = s_view[v];
        q.px = sm.px; q.py = sm.py; q.cz = hdot(Hm + 8, x, y, z); q.d = sm.d; q.weight = sm.weight; q.vis = sm.vis ? 1 : 0;
        // footprint with explicit in-range flags (the backward needs them: a zero weight can also be a cell-border weight)
        const float ix = to_map_index<RECIP>(sm.px, sp.W, sp.w), iy = to_map_index<RECIP>(sm.py, sp.H, sp.h);
        const float x0 = floorf(ix), y0 = floorf(iy);
        const float wx = __fsub_rn(ix, x0), wy = __fsub_rn(iy, y0);
        const float ex = __fsub_rn(1.f, wx), ey = __fsub_rn(1.f, wy);
        const float x1 = x0 + 1.f, y1 = y0 + 1.f, xm = (float)(sp.w - 1), ym = (float)(sp.h - 1);
        const bool x0ok = x0 >= 0.f && x0 <= xm, x1ok = x1 >= 0.f && x1 <= xm;
        const bool y0ok = y0 >= 0.f && y0 <= ym, y1ok = y1 >= 0.f && y1 <= ym;
        const int x0c = (int)fminf(fmaxf(x0, 0.f), xm), x1c = (int)fminf(fmaxf(x1, 0.f), xm);
        const int y0c = (int)fminf(fmaxf(y0, 0.f), ym), y1c = (int)fminf(fmaxf(y1, 0.f), ym);
        q.wx = wx; q.wy = wy;
        q.m[0] = (x0ok && y0ok) ? 1.f : 0.f; q.m[1] = (x1ok && y0ok) ? 1.f : 0.f;
        q.m[2] = (x0ok && y1ok) ? 1.f : 0.f; q.m[3] = (x1ok && y1ok) ? 1.f : 0.f;
        q.fw[0] = (x0ok && y0ok) ? __fmul_rn(ey, ex) : 0.f; q.fw[1] = (x1ok && y0ok) ? __fmul_rn(ey, wx) : 0.f;   // footprint()
        q.fw[2] = (x0ok && y1ok) ? __fmul_rn(wy, ex) : 0.f; q.fw[3] = (x1ok && y1ok) ? __fmul_rn(wy, wx) : 0.f;
        q.o[0] = y0c * sp.sy + x0c * sp.sx; q.o[1] = y0c * sp.sy + x1c * sp.sx;
        q.o[2] = y1c * sp.sy + x0c * sp.sx; q.o[3] = y1c * sp.sy + x1c * sp.sx;
    }
    __syncthreads();

    // dist / valid_mask as d3f_eval: visible views in order (fusion.py:343-370)
    float acc = 0.f, cnt = 0.f;
    for (int v = 0; v < V; ++v)
        if (s_view[v].vis) {
            acc = __fadd_rn(acc, fminf(fmaxf(s_view[v].d, -sp.mu), sp.mu));
            cnt = __fadd_rn(cnt, 1.f);
        }
    const float denom = __fadd_rn(cnt, 1e-6f);
    const bool valid = cnt != 0.f;
    const float dist = valid ? __fdiv_rn(acc, denom) : 1e3f;
    const float inv = __fdiv_rn(1.f, denom);

    // the texel rows of every visible view: 16 independent 128-bit loads per thread
    float4 tex[STEP_VIEWS][4];
#pragma unroll
    for (int v = 0; v < STEP_VIEWS; ++v) {
        const bool on = v < V && s_view[v].vis && active;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            tex[v][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on) tex[v][k] = __ldg(reinterpret_cast<const float4*>(sp.map + (size_t)v * (size_t)sp.sv + s_view[v].o[k] + c));
        }
    }
    // forward: the descriptor of the point, d3f_eval's arithmetic (folded weights, views in order, corners nw ne sw se)
    float4 feat = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < STEP_VIEWS; ++v) {
        if (!(v < V && s_view[v].vis)) continue;
        const float fac = __fmul_rn(s_view[v].weight, inv);                     // fusion.py:385
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float wk = __fmul_rn(s_view[v].fw[k], fac);
            feat.x = fmaf(wk, tex[v][k].x, feat.x); feat.y = fmaf(wk, tex[v][k].y, feat.y);
            feat.z = fmaf(wk, tex[v][k].z, feat.z); feat.w = fmaf(wk, tex[v][k].w, feat.w);
        }
    }
    // loss = mean_p(|feat_p - src_p|_2 valid_p) + dist_w mean_p(max(dist_p valid_p, 0))   (fusion.py:1651-1653)
    float4 df = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) df = make_float4(feat.x - s4.x, feat.y - s4.y, feat.z - s4.z, feat.w - s4.w);
    float ss = fmaf(df.x, df.x, fmaf(df.y, df.y, fmaf(df.z, df.z, df.w * df.w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_ss[warp] = ss;
    __syncthreads();
    ss = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < STEP_THREADS / 32; ++w8) ss += s_ss[w8];
    const float nrm = sqrtf(ss);
    const float vf = valid ? 1.f : 0.f;
    const float scale = (nrm > 0.f) ? vf / (nrm * (float)n) : 0.f;              // torch's norm backward: diff/|diff| (0 at 0)
    const float4 g = make_float4(df.x * scale, df.y * scale, df.z * scale, df.w * scale);

    // backward through the bilinear samples and the weights, from the registers (d3f_backward.cuh: view_gradient)
    float part[STEP_VIEWS * 3];
#pragma unroll
    for (int v = 0; v < STEP_VIEWS; ++v) {
        float s_ix = 0.f, s_iy = 0.f, s_r = 0.f;
        if (v < V && s_view[v].vis) {
            const float wx = s_view[v].wx, wy = s_view[v].wy, ex = 1.f - wx, ey = 1.f - wy;
            const float m00 = s_view[v].m[0], m01 = s_view[v].m[1], m10 = s_view[v].m[2], m11 = s_view[v].m[3];
#define D3F_STEP_LANE(q)                                                                                                 \
            {                                                                                                            \
                const float a = tex[v][0].q * m00, b = tex[v][1].q * m01, cc = tex[v][2].q * m10, d2 = tex[v][3].q * m11; \
                s_ix = fmaf(g.q, (b - a) * ey + (d2 - cc) * wy, s_ix);                                                   \
                s_iy = fmaf(g.q, (cc - a) * ex + (d2 - b) * wx, s_iy);                                                   \
                s_r = fmaf(g.q, (a * ex + b * wx) * ey + (cc * ex + d2 * wx) * wy, s_r);                                 \
            }
            D3F_STEP_LANE(x) D3F_STEP_LANE(y) D3F_STEP_LANE(z) D3F_STEP_LANE(w)
#undef D3F_STEP_LANE
        }
        part[v * 3 + 0] = s_ix; part[v * 3 + 1] = s_iy; part[v * 3 + 2] = s_r;
    }
#pragma unroll
    for (int k = 0; k < STEP_VIEWS * 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
        if (lane == 0) s_red[warp][k] = part[k];
    }
    __syncthreads();

    if (tid < V) {                                       // chain view v to the camera frame and to the point
        const int v = tid;
        float c3[3] = {0.f, 0.f, 0.f};
        if (s_view[v].vis) {
            float s_ix = 0.f, s_iy = 0.f, s_r = 0.f;
            for (int w8 = 0; w8 < STEP_THREADS / 32; ++w8) {
                s_ix += s_red[w8][v * 3 + 0]; s_iy += s_red[w8][v * 3 + 1]; s_r += s_red[w8][v * 3 + 2];
            }
            const float px = s_view[v].px, py = s_view[v].py, cz = s_view[v].cz, d = s_view[v].d, weight = s_view[v].weight;
            const float fac = weight * inv;
            const float G_px = fac * s_ix * ((float)(sp.w - 1) / (float)(sp.W - 1));
            const float G_py = fac * s_iy * ((float)(sp.h - 1) / (float)(sp.H - 1));
            const float G_w = inv * s_r;
            // d loss / d dist: torch's clamp(min=0) passes the gradient where x >= 0
            const float gd = (dist * vf >= 0.f) ? sp.dist_w * vf / (float)n : 0.f;
            float G_d = (fabsf(d) >= sp.mu) ? G_w * weight * (d > 0.f ? -1.f : 1.f) / sp.mu : 0.f;
            if (d >= -sp.mu && d <= sp.mu) G_d += gd * inv;
            const float G_cx = G_px / cz, G_cy = G_py / cz;
            const float G_cz = -G_d - (G_px * px + G_py * py) / cz;
            const float* Hm = sH + v * 12;
            c3[0] = Hm[0] * G_cx + Hm[4] * G_cy + Hm[8] * G_cz;
            c3[1] = Hm[1] * G_cx + Hm[5] * G_cy + Hm[9] * G_cz;
            c3[2] = Hm[2] * G_cx + Hm[6] * G_cy + Hm[10] * G_cz;
        }
        s_c[v][0] = c3[0]; s_c[v][1] = c3[1]; s_c[v][2] = c3[2];
    }
    __syncthreads();
    if (tid < 3) {
        float gp = 0.f;
        for (int v = 0; v < V; ++v)
            if (s_view[v].vis) gp += s_c[v][tid];
        __stcg(sp.grad_pts + (size_t)i * 3 + tid, gp);
        if (tp.pts) tp.pts[(size_t)i * 3 + tid] = tid == 0 ? x : (tid == 1 ? y : z);
    }
    if (tid == 3 && sp.loss_terms) sp.loss_terms[i] = (nrm * vf + sp.dist_w * fmaxf(dist * vf, 0.f)) / (float)n;

    // the last CTA of the instance to arrive takes the Adam step for it
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned arrived = atomicAdd(sp.arrivals + inst, 1u);
        s_last = arrived == (unsigned)tp.P - 1u;
        if (s_last) { sp.arrivals[inst] = 0u; __threadfence(); }          // the next launch starts from zero
    }
    __syncthreads();
    if (!s_last) return;
    TrackParams up = tp;
    up.grad_pts = sp.grad_pts;
    track_update_instance<STEP_THREADS, true>(up, inst, s_red, sR, sT);
}

}  // namespace d3f
