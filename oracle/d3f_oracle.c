/*
 * TEST INFRASTRUCTURE — plain-C restatement of the reference's field query (not product code).
 *
 * Same algorithm and the same float32 operation order as oracle/field_oracle.py, which is pinned
 * bit-for-bit (dist, valid_mask) against golden vectors produced by the unmodified reference
 * (tests/golden/, oracle/gen_golden.py).  This copy exists so the checker can cover BASELINE.json's
 * full sizes (1M points x 1024 channels) in seconds on the GPU box's host cores; tests compare it with
 * the numpy restatement and the golden vectors before it is trusted (tests/test_oracle_golden.py).
 *
 * Reference lines followed (paths relative to /root/reference):
 *   oracle_project_row / hdot   fusion.py:32-55   project_points_coords (small-matrix bmm: k sequential, mul then add)
 *   to_map_index                fusion.py:72-73 + torch CPU grid_sample un-normalise, align_corners=True
 *   nearest depth               fusion.py:327-333 (round-half-to-even, zero padding)
 *   visibility / weight         fusion.py:343-347 (eval), :419-423 (eval_dist)
 *   view reduction              fusion.py:358-370, :385-386
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every multiply and add is rounded separately).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORACLE_MAX_VIEWS 16

static float to_map_index(float p, int img_size, int map_size) {
    float n = p / (float)(img_size - 1);
    n = n * 2.0f;
    n = n - 1.0f;
    float sf = (float)(map_size - 1) / 2.0f;
    return (n + 1.0f) * sf;
}

static float hdot(const float* r, float x, float y, float z) {
    float a = 0.0f + r[0] * x;
    a = a + r[1] * y;
    a = a + r[2] * z;
    a = a + r[3] * 1.0f;
    return a;
}

/* maps[k]: (V,h,w,C) float32 (dtype 0) or uint8 (dtype 1); out[k]: (n,C); inter[k]: (V,n,C) or NULL */
typedef struct {
    int V, H, W;
    const float* depth;
    const float* pts;
    int64_t n, i0, i1;
    int n_keys;
    const void* const* maps;
    const int *dtype, *mh, *mw, *mC;
    float mu;
    int eval_dist;
    float* dist;
    uint8_t* valid;
    float* const* out;
    float* const* inter;
    float Hm[ORACLE_MAX_VIEWS][12];
} Job;

static void* run_range(void* arg) {
    const Job* jb = (const Job*)arg;
    const int V = jb->V, H = jb->H, W = jb->W, n_keys = jb->n_keys, eval_dist = jb->eval_dist;
    const float* depth = jb->depth;
    const float* pts = jb->pts;
    const int64_t n = jb->n;
    const void* const* maps = jb->maps;
    const int *dtype = jb->dtype, *mh = jb->mh, *mw = jb->mw, *mC = jb->mC;
    const float mu = jb->mu;
    float* dist = jb->dist;
    uint8_t* valid = jb->valid;
    float* const* out = jb->out;
    float* const* inter = jb->inter;
    const float (*Hm)[12] = jb->Hm;
    for (int64_t i = jb->i0; i < jb->i1; ++i) {
        const float x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
        float px[ORACLE_MAX_VIEWS], py[ORACLE_MAX_VIEWS], wt[ORACLE_MAX_VIEWS], dcl[ORACLE_MAX_VIEWS];
        int vis[ORACLE_MAX_VIEWS];
        float acc = 0.0f, cnt = 0.0f;
        for (int v = 0; v < V; ++v) {
            float cx = hdot(Hm[v], x, y, z), cy = hdot(Hm[v] + 4, x, y, z), cz = hdot(Hm[v] + 8, x, y, z);
            int ok = !(fabsf(cz) < 1e-4f);
            if (!ok) cz = 1e-3f;
            px[v] = cx / cz;
            py[v] = cy / cz;
            float xr = nearbyintf(to_map_index(px[v], W, W));
            float yr = nearbyintf(to_map_index(py[v], H, H));
            float dep = 0.0f;
            if (xr >= 0.0f && xr <= (float)(W - 1) && yr >= 0.0f && yr <= (float)(H - 1))
                dep = depth[((size_t)v * H + (size_t)(int)yr) * W + (size_t)(int)xr];
            float d = dep - cz;
            if (eval_dist) {
                vis[v] = (dep > 0.0f) && ok;
                wt[v] = 1.0f;
                dcl[v] = d;
            } else {
                vis[v] = (dep > 0.0f) && ok && (d > -mu);
                float a = fminf(mu - fabsf(d), 0.0f);
                wt[v] = expf(a / mu);
                dcl[v] = fminf(fmaxf(d, -mu), mu);
            }
            /* sum(0) over views, in order; an invisible view adds d*0 */
            acc = acc + (vis[v] ? dcl[v] : dcl[v] * 0.0f);
            cnt = cnt + (vis[v] ? 1.0f : 0.0f);
        }
        const float denom = cnt + 1e-6f;
        float dv = acc / denom;
        const int none = (cnt == 0.0f);
        if (!eval_dist && none) dv = 1e3f;
        dist[i] = dv;
        valid[i] = none ? 0 : 1;
        if (eval_dist) continue;
        for (int k = 0; k < n_keys; ++k) {
            const int h = mh[k], w = mw[k], C = mC[k];
            float* o = out[k] + (size_t)i * C;
            for (int c = 0; c < C; ++c) o[c] = 0.0f;
            for (int v = 0; v < V; ++v) {
                float ix = to_map_index(px[v], W, w), iy = to_map_index(py[v], H, h);
                float x0 = floorf(ix), y0 = floorf(iy);
                float wx = ix - x0, ex = 1.0f - wx, wy = iy - y0, sy = 1.0f - wy;
                float cw[4] = {sy * ex, sy * wx, wy * ex, wy * wx};
                float cxs[4] = {x0, x0 + 1.0f, x0, x0 + 1.0f}, cys[4] = {y0, y0, y0 + 1.0f, y0 + 1.0f};
                float* it = (inter && inter[k]) ? inter[k] + ((size_t)v * n + i) * C : NULL;
                const float visf = vis[v] ? 1.0f : 0.0f;
                for (int c = 0; c < C; ++c) {
                    float s = 0.0f;
                    for (int q = 0; q < 4; ++q) {
                        float val = 0.0f;
                        if (cxs[q] >= 0.0f && cxs[q] <= (float)(w - 1) && cys[q] >= 0.0f && cys[q] <= (float)(h - 1)) {
                            size_t t = (((size_t)v * h + (size_t)(int)cys[q]) * w + (size_t)(int)cxs[q]) * C + c;
                            val = dtype[k] == 0 ? ((const float*)maps[k])[t] : (float)((const uint8_t*)maps[k])[t];
                        }
                        s = s + val * cw[q];
                    }
                    if (it) it[c] = s;
                    o[c] = o[c] + (s * visf) * wt[v];
                }
            }
            for (int c = 0; c < C; ++c) o[c] = none ? 0.0f : o[c] / denom;
        }
    }
    return NULL;
}

static int g_threads = 0;

int d3f_oracle_threads(void) {
    if (g_threads > 0) return g_threads;
    long t = sysconf(_SC_NPROCESSORS_ONLN);
    return t < 1 ? 1 : (t > 256 ? 256 : (int)t);
}

void d3f_oracle_set_threads(int t) { g_threads = t; }

int d3f_oracle_eval(int V, int H, int W, const float* pose, const float* K, const float* depth,
                    const float* pts, int64_t n, int n_keys, const void* const* maps, const int* dtype,
                    const int* mh, const int* mw, const int* mC, float mu, int eval_dist,
                    float* dist, uint8_t* valid, float* const* out, float* const* inter) {
    if (V < 1 || V > ORACLE_MAX_VIEWS) return -1;
    Job base;
    memset(&base, 0, sizeof(base));
    base.V = V; base.H = H; base.W = W; base.depth = depth; base.pts = pts; base.n = n; base.n_keys = n_keys;
    base.maps = maps; base.dtype = dtype; base.mh = mh; base.mw = mw; base.mC = mC; base.mu = mu;
    base.eval_dist = eval_dist; base.dist = dist; base.valid = valid; base.out = out; base.inter = inter;
    for (int v = 0; v < V; ++v)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) {
                float a = 0.0f;
                for (int k = 0; k < 3; ++k) a = a + K[v * 9 + i * 3 + k] * pose[v * 12 + k * 4 + j];
                base.Hm[v][i * 4 + j] = a;
            }
    int T = d3f_oracle_threads();
    if ((int64_t)T > (n + 255) / 256) T = (int)((n + 255) / 256);
    if (T <= 1) {
        base.i0 = 0; base.i1 = n;
        run_range(&base);
        return 0;
    }
    Job* jobs = (Job*)malloc(sizeof(Job) * (size_t)T);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)T);
    if (!jobs || !th) { free(jobs); free(th); return -2; }
    for (int t = 0; t < T; ++t) {
        jobs[t] = base;
        jobs[t].i0 = n * t / T;
        jobs[t].i1 = n * (t + 1) / T;
        if (pthread_create(&th[t], NULL, run_range, &jobs[t]) != 0) { run_range(&jobs[t]); th[t] = 0; }
    }
    for (int t = 0; t < T; ++t) if (th[t]) pthread_join(th[t], NULL);
    free(jobs); free(th);
    return 0;
}
