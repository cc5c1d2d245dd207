/*
 * lrc_oracle.c -- CPU ORACLE for the LiDAR ray-casting hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (the CUDA engine under the package's csrc/) never links,
 * imports or calls anything in oracle/.
 *
 * What it restates (citations are relative to the reference checkout, /root/reference):
 *   - lidar/indoor_lidar.py:94-131   single-axis ray table (vertical_degrees path)
 *   - lidar/indoor_lidar.py:56-91    single-axis ray table (uniform fov path)
 *   - lidar/indoor_lidar.py:224-296  dual-axis swinging-line ray table, angle noise, dropout
 *   - raycast_engine/raycast_engine_cpu.py:46-62   closest-hit cast + f32 point reconstruction
 *   - raycast_engine/raycast_engine_cpu.py:95-107  f64 range filter + incident angle
 * The intersector itself lives in a third-party dependency that is NOT in the reference tree
 * (open3d>=0.17.0, requirements.txt:2, un-pinned; it wraps Intel Embree).  Its published
 * contract -- closest hit, t in [0, inf), two-sided triangles, vertices rounded to float32,
 * t measured in units of |d| -- is restated here with a Moller-Trumbore test whose float32
 * operation order is fixed (see mt_f32 below) so that the CUDA engine can be compared bit for
 * bit.  PARITY STATUS: ray tables are pinned by golden vectors produced from the live
 * reference `lidar` package (tests/golden/); the intersector is "parity unpinned" against
 * Open3D/Embree (absent, uninstallable) and pinned instead by analytic known answers and a
 * float64 brute-force cross-check (tests/test_oracle_*.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -mavx2 -mfma -fopenmp).
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is an explicit fmaf()/fma().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))
#define ORC_MISS_ID 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------
 * Deterministic float32 Moller-Trumbore ("division deferred" form).
 *   p   = d x e2                      each component  fmaf(a,b, -(c*e))
 *   det = e1 . p                      fmaf(x,x', fmaf(y,y', z*z'))
 *   tv  = o - v0
 *   U   = tv . p ;  q = tv x e1 ;  V = d . q ;  W = e2 . q
 *   if det < 0: negate det, U, V, W   (exact)
 *   hit  <=>  det > 0  &&  U >= 0  &&  V >= 0  &&  (U + V) <= det  &&  W >= 0
 *   t = W / det                       (IEEE division)
 * Two-sided (no back-face culling), t >= 0 accepted, t in units of |d|.
 * ------------------------------------------------------------------------------------------ */
static inline float dot3_f32(float ax, float ay, float az, float bx, float by, float bz)
{
    return fmaf(ax, bx, fmaf(ay, by, az * bz));
}

static inline int mt_f32(const float o[3], const float d[3], const float v0[3], const float e1[3],
                         const float e2[3], float *t_out)
{
    float px = fmaf(d[1], e2[2], -(d[2] * e2[1]));
    float py = fmaf(d[2], e2[0], -(d[0] * e2[2]));
    float pz = fmaf(d[0], e2[1], -(d[1] * e2[0]));
    float det = dot3_f32(e1[0], e1[1], e1[2], px, py, pz);
    float tx = o[0] - v0[0], ty = o[1] - v0[1], tz = o[2] - v0[2];
    float U = dot3_f32(tx, ty, tz, px, py, pz);
    float qx = fmaf(ty, e1[2], -(tz * e1[1]));
    float qy = fmaf(tz, e1[0], -(tx * e1[2]));
    float qz = fmaf(tx, e1[1], -(ty * e1[0]));
    float V = dot3_f32(d[0], d[1], d[2], qx, qy, qz);
    float W = dot3_f32(e2[0], e2[1], e2[2], qx, qy, qz);
    if (det < 0.0f) { det = -det; U = -U; V = -V; W = -W; }
    if (det > 0.0f && U >= 0.0f && V >= 0.0f && (U + V) <= det && W >= 0.0f) {
        float t = W / det;
        if (t < INFINITY) { *t_out = t; return 1; }
    }
    return 0;
}

/* Same formulas evaluated in float64 on the float32 inputs: the tolerance anchor. */
static inline int mt_f64(const float o[3], const float d[3], const float v0[3], const float e1f[3],
                         const float e2f[3], double *t_out)
{
    double e1[3] = {e1f[0], e1f[1], e1f[2]}, e2[3] = {e2f[0], e2f[1], e2f[2]};
    double dx = d[0], dy = d[1], dz = d[2];
    double px = dy * e2[2] - dz * e2[1], py = dz * e2[0] - dx * e2[2], pz = dx * e2[1] - dy * e2[0];
    double det = e1[0] * px + e1[1] * py + e1[2] * pz;
    double tx = (double)o[0] - v0[0], ty = (double)o[1] - v0[1], tz = (double)o[2] - v0[2];
    double U = tx * px + ty * py + tz * pz;
    double qx = ty * e1[2] - tz * e1[1], qy = tz * e1[0] - tx * e1[2], qz = tx * e1[1] - ty * e1[0];
    double V = dx * qx + dy * qy + dz * qz;
    double W = e2[0] * qx + e2[1] * qy + e2[2] * qz;
    if (det < 0.0) { det = -det; U = -U; V = -V; W = -W; }
    if (det > 0.0 && U >= 0.0 && V >= 0.0 && (U + V) <= det && W >= 0.0) {
        *t_out = W / det;
        return 1;
    }
    return 0;
}

/* closest-hit bookkeeping: smallest t, ties -> smallest original triangle id */
static inline void take_hit(float t, uint32_t id, float *best_t, uint32_t *best_id)
{
    if (t < *best_t || (t == *best_t && id < *best_id)) { *best_t = t; *best_id = id; }
}

/* ------------------------------------------------------------------------------------------
 * Scene: float32 triangles as (v0, e1 = v1 - v0, e2 = v2 - v0) + a binned-SAH BVH2.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float v0[3], e1[3], e2[3]; } orc_tri;

typedef struct {
    float lo[3], hi[3];
    int32_t left;   /* internal: index of left child; leaf: first slot in `order` */
    int32_t right;  /* internal: index of right child; leaf: -(count) */
} orc_node;

typedef struct orc_scene {
    int64_t T;
    orc_tri *tri;       /* original order */
    float *blo, *bhi;   /* per-triangle padded boxes, T*3 each */
    float *cen;         /* centroids T*3 */
    int32_t *order;     /* triangle ids in leaf order */
    orc_node *node;
    int32_t n_nodes;
    int32_t root;
    float pad;
    /* statistics of the last orc_cast_rays call */
    uint64_t stat_nodes, stat_tris, stat_rays;
} orc_scene;

static void tri_setup(const float *verts, const int32_t *idx, int64_t T, orc_tri *tri)
{
#pragma omp parallel for schedule(static) if (T >= 131072)
    for (int64_t i = 0; i < T; ++i) {
        const float *a = verts + 3 * (int64_t)idx[3 * i + 0];
        const float *b = verts + 3 * (int64_t)idx[3 * i + 1];
        const float *c = verts + 3 * (int64_t)idx[3 * i + 2];
        for (int k = 0; k < 3; ++k) {
            tri[i].v0[k] = a[k];
            tri[i].e1[k] = b[k] - a[k];
            tri[i].e2[k] = c[k] - a[k];
        }
    }
}

#define NBINS 16
#define LEAF_MAX 4

typedef struct { float lo[3], hi[3]; } box3;
static inline void box_empty(box3 *b)
{
    for (int k = 0; k < 3; ++k) { b->lo[k] = INFINITY; b->hi[k] = -INFINITY; }
}
static inline void box_grow(box3 *b, const float *lo, const float *hi)
{
    for (int k = 0; k < 3; ++k) {
        if (lo[k] < b->lo[k]) b->lo[k] = lo[k];
        if (hi[k] > b->hi[k]) b->hi[k] = hi[k];
    }
}
static inline float box_area(const box3 *b)
{
    float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

static int32_t node_alloc(orc_scene *s)
{
    int32_t i;
#pragma omp atomic capture
    i = s->n_nodes++;
    return i;
}

static int32_t build_rec(orc_scene *s, int32_t first, int32_t count, int depth)
{
    int32_t me = node_alloc(s);
    orc_node *n = &s->node[me];
    box3 bb, cb;
    box_empty(&bb);
    box_empty(&cb);
    for (int32_t i = first; i < first + count; ++i) {
        int32_t id = s->order[i];
        box_grow(&bb, s->blo + 3 * (int64_t)id, s->bhi + 3 * (int64_t)id);
        box_grow(&cb, s->cen + 3 * (int64_t)id, s->cen + 3 * (int64_t)id);
    }
    memcpy(n->lo, bb.lo, sizeof bb.lo);
    memcpy(n->hi, bb.hi, sizeof bb.hi);
    if (count <= LEAF_MAX) { n->left = first; n->right = -count; return me; }

    /* binned SAH over the axis of largest centroid extent */
    int axis = 0;
    float ext = cb.hi[0] - cb.lo[0];
    for (int k = 1; k < 3; ++k)
        if (cb.hi[k] - cb.lo[k] > ext) { ext = cb.hi[k] - cb.lo[k]; axis = k; }
    int32_t mid = first + count / 2;
    if (ext > 0.0f) {
        box3 bin_box[NBINS];
        int32_t bin_cnt[NBINS];
        for (int b = 0; b < NBINS; ++b) { box_empty(&bin_box[b]); bin_cnt[b] = 0; }
        float scale = (float)NBINS * (1.0f - 1e-6f) / ext;
        for (int32_t i = first; i < first + count; ++i) {
            int32_t id = s->order[i];
            int b = (int)((s->cen[3 * (int64_t)id + axis] - cb.lo[axis]) * scale);
            if (b < 0) b = 0;
            if (b >= NBINS) b = NBINS - 1;
            bin_cnt[b]++;
            box_grow(&bin_box[b], s->blo + 3 * (int64_t)id, s->bhi + 3 * (int64_t)id);
        }
        float right_area[NBINS];
        int32_t right_cnt[NBINS];
        box3 acc;
        box_empty(&acc);
        int32_t c = 0;
        for (int b = NBINS - 1; b > 0; --b) {
            box_grow(&acc, bin_box[b].lo, bin_box[b].hi);
            c += bin_cnt[b];
            right_area[b] = box_area(&acc);
            right_cnt[b] = c;
        }
        box_empty(&acc);
        c = 0;
        float best = INFINITY;
        int best_b = -1;
        for (int b = 0; b < NBINS - 1; ++b) {
            box_grow(&acc, bin_box[b].lo, bin_box[b].hi);
            c += bin_cnt[b];
            if (c == 0 || right_cnt[b + 1] == 0) continue;
            float cost = box_area(&acc) * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
            if (cost < best) { best = cost; best_b = b; }
        }
        if (best_b >= 0) {
            int32_t i = first, j = first + count - 1;
            while (i <= j) {
                int32_t id = s->order[i];
                int b = (int)((s->cen[3 * (int64_t)id + axis] - cb.lo[axis]) * scale);
                if (b < 0) b = 0;
                if (b >= NBINS) b = NBINS - 1;
                if (b <= best_b) ++i;
                else { s->order[i] = s->order[j]; s->order[j] = id; --j; }
            }
            if (i > first && i < first + count) mid = i;
        }
    }
    int32_t l, r;
    if (count > 4096 && depth < 24) {
#pragma omp task shared(l) firstprivate(first, mid, depth)
        l = build_rec(s, first, mid - first, depth + 1);
#pragma omp task shared(r) firstprivate(first, mid, count, depth)
        r = build_rec(s, mid, first + count - mid, depth + 1);
#pragma omp taskwait
    } else {
        l = build_rec(s, first, mid - first, depth + 1);
        r = build_rec(s, mid, first + count - mid, depth + 1);
    }
    s->node[me].left = l;
    s->node[me].right = r;
    return me;
}

ORC_API orc_scene *orc_scene_create(const float *verts, int64_t V, const int32_t *idx, int64_t T)
{
    (void)V;
    orc_scene *s = (orc_scene *)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->T = T;
    if (T == 0) return s;
    s->tri = (orc_tri *)malloc(sizeof(orc_tri) * T);
    s->blo = (float *)malloc(sizeof(float) * 3 * T);
    s->bhi = (float *)malloc(sizeof(float) * 3 * T);
    s->cen = (float *)malloc(sizeof(float) * 3 * T);
    s->order = (int32_t *)malloc(sizeof(int32_t) * T);
    s->node = (orc_node *)malloc(sizeof(orc_node) * (2 * T + 1));
    tri_setup(verts, idx, T, s->tri);
    /* scene extent -> padding so that every accepted Moller-Trumbore hit lies inside its box */
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < T; ++i)
        for (int c = 0; c < 3; ++c) {
            const float *p = verts + 3 * (int64_t)idx[3 * i + c];
            for (int k = 0; k < 3; ++k) {
                if (p[k] < lo[k]) lo[k] = p[k];
                if (p[k] > hi[k]) hi[k] = p[k];
            }
        }
    float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
    float amax = 0.0f;
    for (int k = 0; k < 3; ++k) amax = fmaxf(amax, fmaxf(fabsf(lo[k]), fabsf(hi[k])));
    s->pad = ldexpf(fmaxf(ext, amax), -17);
#pragma omp parallel for schedule(static) if (T >= 131072)
    for (int64_t i = 0; i < T; ++i) {
        for (int k = 0; k < 3; ++k) {
            float a = verts[3 * (int64_t)idx[3 * i + 0] + k];
            float b = verts[3 * (int64_t)idx[3 * i + 1] + k];
            float c = verts[3 * (int64_t)idx[3 * i + 2] + k];
            float mn = fminf(a, fminf(b, c)), mx = fmaxf(a, fmaxf(b, c));
            s->blo[3 * i + k] = mn - s->pad;
            s->bhi[3 * i + k] = mx + s->pad;
            s->cen[3 * i + k] = 0.5f * (mn + mx);
        }
        s->order[i] = (int32_t)i;
    }
    s->n_nodes = 0;
    if (T < 131072) {
        /* small scenes: task start-up costs more than it buys */
        s->root = build_rec(s, 0, (int32_t)T, 0);
    } else {
#pragma omp parallel
#pragma omp single
        s->root = build_rec(s, 0, (int32_t)T, 0);
    }
    return s;
}

ORC_API void orc_scene_destroy(orc_scene *s)
{
    if (!s) return;
    free(s->tri); free(s->blo); free(s->bhi); free(s->cen); free(s->order); free(s->node);
    free(s);
}

static inline float safe_inv(float d)
{
    const float eps = 1e-20f;
    if (fabsf(d) < eps) d = copysignf(eps, d);
    return 1.0f / d;
}

#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))
static inline int slab(const orc_node *n, const float o[3], const float inv[3], float tmax, float *tnear)
{
    float t0 = 0.0f, t1 = tmax;
    for (int k = 0; k < 3; ++k) {
        float a = (n->lo[k] - o[k]) * inv[k];
        float b = (n->hi[k] - o[k]) * inv[k];
        float mn = ORC_MIN(a, b), mx = ORC_MAX(a, b);
        /* widen by a few ulps: the test must never reject a box the exact ray enters */
        mn -= fabsf(mn) * 4e-7f;
        mx += fabsf(mx) * 4e-7f;
        t0 = ORC_MAX(t0, mn);
        t1 = ORC_MIN(t1, mx);
    }
    *tnear = t0;
    return t0 <= t1;
}

static void cast_one(const orc_scene *s, const float *ray, float *t_out, uint32_t *id_out,
                     uint64_t *n_nodes, uint64_t *n_tris)
{
    float best_t = INFINITY;
    uint32_t best_id = ORC_MISS_ID;
    const float *o = ray, *d = ray + 3;
    if (s->T > 0) {
        float inv[3] = {safe_inv(d[0]), safe_inv(d[1]), safe_inv(d[2])};
        struct { int32_t node; float tnear; } stack[128];
        int sp = 0, overflow = 0;
        float tn;
        ++*n_nodes;
        if (slab(&s->node[s->root], o, inv, best_t, &tn)) { stack[0].node = s->root; stack[0].tnear = tn; sp = 1; }
        while (sp > 0) {
            --sp;
            if (stack[sp].tnear > best_t) continue;   /* equal t may still hide a smaller id: visit */
            const orc_node *n = &s->node[stack[sp].node];
            if (n->right < 0) {
                int32_t cnt = -n->right;
                for (int32_t i = 0; i < cnt; ++i) {
                    int32_t id = s->order[n->left + i];
                    const orc_tri *tr = &s->tri[id];
                    float t;
                    ++*n_tris;
                    if (mt_f32(o, d, tr->v0, tr->e1, tr->e2, &t)) take_hit(t, (uint32_t)id, &best_t, &best_id);
                }
                continue;
            }
            float tl, tr;
            *n_nodes += 2;
            int hl = slab(&s->node[n->left], o, inv, best_t, &tl);
            int hr = slab(&s->node[n->right], o, inv, best_t, &tr);
            if (sp + 2 > 128) { overflow = 1; break; }
            /* far child first so that the near one is popped next */
            if (hl && hr) {
                int near_left = tl <= tr;
                stack[sp].node = near_left ? n->right : n->left; stack[sp++].tnear = near_left ? tr : tl;
                stack[sp].node = near_left ? n->left : n->right; stack[sp++].tnear = near_left ? tl : tr;
            } else if (hl) { stack[sp].node = n->left; stack[sp++].tnear = tl; }
            else if (hr) { stack[sp].node = n->right; stack[sp++].tnear = tr; }
        }
        if (overflow) { /* pathological depth: exhaustive scan gives the same answer */
            best_t = INFINITY; best_id = ORC_MISS_ID;
            for (int64_t i = 0; i < s->T; ++i) {
                float t;
                if (mt_f32(o, d, s->tri[i].v0, s->tri[i].e1, s->tri[i].e2, &t))
                    take_hit(t, (uint32_t)i, &best_t, &best_id);
            }
        }
    }
    *t_out = best_t;
    *id_out = best_id;
}

/* rays: N x 6 float32 (origin, direction as given -- NOT normalised, cf. raycast_engine_cpu.py:50-51) */
ORC_API int orc_cast_rays(orc_scene *s, const float *rays, int64_t N, float *t_hit, uint32_t *prim_id)
{
    uint64_t nn = 0, nt = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nn, nt)
    for (int64_t i = 0; i < N; ++i) cast_one(s, rays + 6 * i, t_hit + i, prim_id + i, &nn, &nt);
    s->stat_nodes = nn; s->stat_tris = nt; s->stat_rays = (uint64_t)N;
    return 0;
}

ORC_API void orc_scene_stats(const orc_scene *s, uint64_t *nodes, uint64_t *tris, uint64_t *rays, int32_t *n_nodes)
{
    *nodes = s->stat_nodes; *tris = s->stat_tris; *rays = s->stat_rays; *n_nodes = s->n_nodes;
}

/* exhaustive float32 closest hit (same mt_f32), no acceleration structure */
ORC_API int orc_cast_rays_brute(const float *verts, const int32_t *idx, int64_t T, const float *rays,
                                int64_t N, float *t_hit, uint32_t *prim_id)
{
    orc_tri *tri = (orc_tri *)malloc(sizeof(orc_tri) * (T > 0 ? T : 1));
    tri_setup(verts, idx, T, tri);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < N; ++r) {
        float best_t = INFINITY;
        uint32_t best_id = ORC_MISS_ID;
        for (int64_t i = 0; i < T; ++i) {
            float t;
            if (mt_f32(rays + 6 * r, rays + 6 * r + 3, tri[i].v0, tri[i].e1, tri[i].e2, &t))
                take_hit(t, (uint32_t)i, &best_t, &best_id);
        }
        t_hit[r] = best_t;
        prim_id[r] = best_id;
    }
    free(tri);
    return 0;
}

/* exhaustive float64 closest hit on the float32 inputs: the tolerance anchor (|dt| <= 1e-4 m) */
ORC_API int orc_cast_rays_brute_f64(const float *verts, const int32_t *idx, int64_t T, const float *rays,
                                    int64_t N, double *t_hit, uint32_t *prim_id)
{
    orc_tri *tri = (orc_tri *)malloc(sizeof(orc_tri) * (T > 0 ? T : 1));
    /* edges in float64 from the float32 vertices */
    double *e = (double *)malloc(sizeof(double) * 6 * (T > 0 ? T : 1));
    tri_setup(verts, idx, T, tri);
    for (int64_t i = 0; i < T; ++i)
        for (int k = 0; k < 3; ++k) {
            double a = verts[3 * (int64_t)idx[3 * i + 0] + k];
            e[6 * i + k] = (double)verts[3 * (int64_t)idx[3 * i + 1] + k] - a;
            e[6 * i + 3 + k] = (double)verts[3 * (int64_t)idx[3 * i + 2] + k] - a;
        }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < N; ++r) {
        const float *o = rays + 6 * r, *d = o + 3;
        double best_t = INFINITY;
        uint32_t best_id = ORC_MISS_ID;
        for (int64_t i = 0; i < T; ++i) {
            const double *e1 = e + 6 * i, *e2 = e1 + 3;
            double dx = d[0], dy = d[1], dz = d[2];
            double px = dy * e2[2] - dz * e2[1], py = dz * e2[0] - dx * e2[2], pz = dx * e2[1] - dy * e2[0];
            double det = e1[0] * px + e1[1] * py + e1[2] * pz;
            double tx = (double)o[0] - tri[i].v0[0], ty = (double)o[1] - tri[i].v0[1], tz = (double)o[2] - tri[i].v0[2];
            double U = tx * px + ty * py + tz * pz;
            double qx = ty * e1[2] - tz * e1[1], qy = tz * e1[0] - tx * e1[2], qz = tx * e1[1] - ty * e1[0];
            double V = dx * qx + dy * qy + dz * qz;
            double W = e2[0] * qx + e2[1] * qy + e2[2] * qz;
            if (det < 0.0) { det = -det; U = -U; V = -V; W = -W; }
            if (det > 0.0 && U >= 0.0 && V >= 0.0 && (U + V) <= det && W >= 0.0) {
                double t = W / det;
                if (t < best_t || (t == best_t && (uint32_t)i < best_id)) { best_t = t; best_id = (uint32_t)i; }
            }
        }
        t_hit[r] = best_t;
        prim_id[r] = best_id;
    }
    free(tri);
    free(e);
    (void)mt_f64;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Ray tables.
 * ------------------------------------------------------------------------------------------ */
static const double ORC_PI = 3.141592653589793;

static inline void rotate_store(const double *pose, double lx, double ly, double lz, float *ray)
{
    /* world = R . local ; origin = pose[:3,3] ; float64 math, one rounding to float32 at the end */
    for (int r = 0; r < 3; ++r) {
        double w = pose[4 * r + 0] * lx + pose[4 * r + 1] * ly + pose[4 * r + 2] * lz;
        ray[r] = (float)pose[4 * r + 3];
        ray[3 + r] = (float)w;
    }
}

/* lidar/indoor_lidar.py:94-131 -- ray index = j*W + i, beta = -(i - W/2)/W*2*pi, alpha = deg2rad(v[j]) */
ORC_API int orc_gen_rays_single_axis(const double *pose, const double *vertical_deg, int H, int W, float *rays)
{
    if (W < 1) W = 1;
    double one_deg = 0.0;
    if (H <= 0 || vertical_deg == NULL) { vertical_deg = &one_deg; H = 1; }   /* :104-106 */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < H; ++j) {
        double alpha = vertical_deg[j] * (ORC_PI / 180.0);
        double ca = cos(alpha), sa = sin(alpha);
        for (int i = 0; i < W; ++i) {
            double beta = -((double)i - (double)W / 2.0) / (double)W * 2.0 * ORC_PI;
            rotate_store(pose, ca * cos(beta), ca * sin(beta), sa, rays + 6 * ((int64_t)j * W + i));
        }
    }
    return 0;
}

/* lidar/indoor_lidar.py:56-91 -- linspace(fov_up, -fov_down, H) x linspace(0, 2pi, W, endpoint=False);
 * local directions are rounded to float32 BEFORE the rotation (:82), the rotation is float64 (:88). */
ORC_API int orc_gen_rays_uniform(const double *pose, double fov_up_deg, double fov_down_deg, int H, int W, float *rays)
{
    if (H < 1) H = 1;
    if (W < 1) W = 1;
    double up = fov_up_deg * (ORC_PI / 180.0), dn = fov_down_deg * (ORC_PI / 180.0);
    double vstep = (H > 1) ? ((-dn) - up) / (double)(H - 1) : 0.0;
    double hstep = (2.0 * ORC_PI - 0.0) / (double)W;
    for (int j = 0; j < H; ++j) {
        double v = (H > 1 && j == H - 1) ? -dn : up + (double)j * vstep;
        double cv = cos(v), sv = sin(v);
        for (int i = 0; i < W; ++i) {
            double h = (double)i * hstep;
            float lx = (float)(cv * cos(h)), ly = (float)(cv * sin(h)), lz = (float)sv;
            rotate_store(pose, (double)lx, (double)ly, (double)lz, rays + 6 * ((int64_t)j * W + i));
        }
    }
    return 0;
}

/* Philox4x32-10 counter-based generator (Salmon et al., SC'11): the engine's noise source.
 * key = (seed_lo, seed_hi); counter = (ray_idx, pose_lo, pose_hi, stream). */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
static inline double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

ORC_API void orc_philox(uint64_t seed, uint64_t pose_idx, uint32_t ray_idx, uint32_t stream, uint32_t out[4])
{
    uint32_t c[4] = {ray_idx, (uint32_t)pose_idx, (uint32_t)(pose_idx >> 32), stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    memcpy(out, c, sizeof c);
}

typedef struct {
    int32_t num_lines;          /* num_vertical_lines */
    int32_t points_per_line;    /* int(point_rate*scan_duration) // num_lines */
    double theta_min, theta_max;/* theta_range */
    double swing_amplitude, swing_frequency;
    double angle_noise_std;     /* rad; 0 disables */
    double dropout_probability; /* 0 disables */
} orc_dual_params;

/* lidar/indoor_lidar.py:224-296.  Rays are written DENSE (line*ppl + k) with keep[] flags; the caller
 * compacts.  Noise: phi += s*z0, theta += s*z1 AFTER the clip (:263-272); dropout keeps u > p (:292-294).
 * The reference draws from the global numpy stream; this engine uses Philox keyed on
 * (seed, pose_idx, ray_idx), so only the noise-free table is bit-comparable with the reference. */
ORC_API int orc_gen_rays_dual_axis(const double *pose, const orc_dual_params *q, uint64_t seed, uint64_t pose_idx,
                                   float *rays, uint8_t *keep, int64_t *n_kept)
{
    int L = q->num_lines, K = q->points_per_line;
    int64_t kept = 0;
    double tstep = (L > 1) ? (q->theta_min - q->theta_max) / (double)(L - 1) : 0.0;
    double pstep = (2.0 * ORC_PI - 0.0) / (double)K;
    for (int line = 0; line < L; ++line) {
        double base = (L > 1 && line == L - 1) ? q->theta_min : q->theta_max + (double)line * tstep;
        double phase = (double)line * ORC_PI / (double)L;
        for (int k = 0; k < K; ++k) {
            int64_t r = (int64_t)line * K + k;
            double phi = (double)k * pstep;
            double swing = q->swing_amplitude * sin(q->swing_frequency * phi + phase);
            double theta = base + swing;
            if (theta < q->theta_min) theta = q->theta_min;
            if (theta > q->theta_max) theta = q->theta_max;
            uint32_t rnd[4] = {0, 0, 0, 0};
            if (q->angle_noise_std > 0.0 || q->dropout_probability > 0.0)
                orc_philox(seed, pose_idx, (uint32_t)r, 0u, rnd);
            if (q->angle_noise_std > 0.0) {
                double rad = sqrt(-2.0 * log(u01(rnd[0])));
                double ang = 2.0 * ORC_PI * u01(rnd[1]);
                phi += q->angle_noise_std * (rad * cos(ang));
                theta += q->angle_noise_std * (rad * sin(ang));
            }
            double ct = cos(theta);
            rotate_store(pose, ct * cos(phi), ct * sin(phi), sin(theta), rays + 6 * r);
            /* :282-283 rounds the origin to float32 exactly as rotate_store does */
            uint8_t kp = 1;
            if (q->dropout_probability > 0.0) kp = (u01(rnd[2]) > q->dropout_probability) ? 1 : 0;
            keep[r] = kp;
            kept += kp;
        }
    }
    *n_kept = kept;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Frame epilogue, raycast_engine_cpu.py:54-62 + :95-107, for a dense (t_hit, prim_id) result.
 *   d^ = d / sqrt((dx*dx + dy*dy) + dz*dz)     float32, separate roundings        (:57)
 *   p  = o + d^ * t                            float32, mul then add              (:62)
 *   keep <=> t != inf  &&  sqrt(((p-c)^2).sum()) < max_range   float64, strict     (:54,:95-97)
 *   incident = degrees(arccos(|(p-c)_z / ||p-c|||))            float64             (:100-107)
 * max_range < 0 disables the range filter (== rays_intersect_mesh alone).
 * Returns the number of kept points; outputs are written compacted in ray order.
 * ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_epilogue(const float *rays, const float *t_hit, const uint32_t *prim_id, const uint8_t *keep_in,
                             int64_t N, const double *center, double max_range, const uint32_t *tri_label,
                             float *xyz, double *incident, uint32_t *out_prim, uint32_t *out_label, uint32_t *out_ray)
{
    int64_t m = 0;
    for (int64_t i = 0; i < N; ++i) {
        if (keep_in && !keep_in[i]) continue;
        float t = t_hit[i];
        if (!(t != INFINITY)) continue;
        const float *o = rays + 6 * i, *d = o + 3;
        float n = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        float p[3];
        for (int k = 0; k < 3; ++k) {
            float dh = d[k] / n;
            float s = dh * t;
            p[k] = o[k] + s;
        }
        double inc = 0.0;
        if (max_range >= 0.0) {
            double dx = (double)p[0] - center[0], dy = (double)p[1] - center[1], dz = (double)p[2] - center[2];
            double dist = sqrt((dx * dx + dy * dy) + dz * dz);
            if (!(dist < max_range)) continue;
            inc = acos(fabs(dz / dist)) * (180.0 / ORC_PI);
        }
        xyz[3 * m + 0] = p[0]; xyz[3 * m + 1] = p[1]; xyz[3 * m + 2] = p[2];
        if (incident) incident[m] = inc;
        if (out_prim) out_prim[m] = prim_id[i];
        if (out_label) out_label[m] = tri_label ? tri_label[prim_id[i]] : 0u;
        if (out_ray) out_ray[m] = (uint32_t)i;
        ++m;
    }
    return m;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
