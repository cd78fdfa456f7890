/*
 * cellflow_oracle.c — CPU restatement of CellFlow's particle-life law.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may build,
 * load or call this file.  The product (cellflow_b200/) never links or imports it; it fails
 * loudly when its CUDA library is missing instead of falling back to anything here.
 *
 * What it restates (all citations relative to /root/reference):
 *   - the step law            cuda-native/src/ParticleSimulation.cu:86-165
 *   - the force-table squash  cuda-native/src/ParticleSimulation.cu:521-527
 *   - the proximity-graph rule cuda-native/src/ParticleSimulation.cu:212-246
 *   - the spawn rule (shape)  cuda-native/src/ParticleSimulation.cu:45-62
 *   - the LFO                 cuda-native/src/CellFlowWidget.cpp:415-421
 *
 * Arithmetic contract.  The reference is a CUDA program, so "the reference's result" is what
 * nvcc makes of those lines.  The rounding points below (which a*b+c are fused, operand order,
 * IEEE sqrt/div) are read off the SASS nvcc 12.9 emits for the unmodified reference file at
 * -arch=sm_100a (see DESIGN.md section 3 for the listing).  This file is compiled with
 * -ffp-contract=off and writes every fused multiply-add as an explicit fmaf(), so gcc cannot
 * move a rounding point.  The one piece a CPU cannot reproduce bit-for-bit is MUFU.EX2 inside
 * CUDA's expf (<= 2 ulp); exp2f from libm stands in for it.  Everything that decides a
 * neighbour (displacement, wrap, distance, radius, comparison) is bit-reproducible, which is
 * what makes neighbour counts and graph edge sets exact.
 *
 * The reference kernel updates particles in place while other threads still read them
 * (.cu:89 vs .cu:164), so its own output depends on scheduling.  The law restated here is the
 * race-free reading of the same code: every particle reads the state of step t and writes
 * the state of step t+1 (Jacobi).
 *
 * Parity pin: the reference ships no tests or golden vectors.  This
 * oracle is pinned against the reference's own kernels compiled from /root/reference and run on
 * a B200 (oracle/ref_harness.cu, launched one warp at a time so that the in-place race cannot
 * occur); the resulting vectors are committed under tests/golden/ with the script that made
 * them.  See tests/test_oracle_golden.py.
 */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cellflow_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Force table: ParticleSimulation.cu:521-527.  In a .cu translation unit the unqualified
 * tanh(float) / fmax / fmin resolve to the float overloads CUDA's headers put in the global
 * namespace, and the host compiler (x86-64 baseline, no FMA) rounds the product and the sum
 * separately.
 * ---------------------------------------------------------------------------------------- */
void orc_force_table(const float* raw, int T, float range, float bias, float offset, float* out) {
    for (int i = 0; i < T * T; i++) {
        float t = tanhf(raw[i] * offset);
        float v = t * range;
        v = v + bias;
        out[i] = fmaxf(-1.0f, fminf(1.0f, v));
    }
}

/* Default tables of a freshly constructed ParticleSimulation: libc rand(), never seeded
 * (.cu:513-519, 533-539).  glibc's rand() with the default seed is a fixed sequence. */
void orc_default_tables(int T, float* raw, float* radio) {
    srand(1); /* the state an unseeded process starts in */
    for (int i = 0; i < T * T; i++) raw[i] = (float)rand() / RAND_MAX * 2.0f - 1.0f;
    for (int i = 0; i < T; i++) radio[i] = (float)rand() / RAND_MAX * 2.0f - 1.0f;
}

/* LFO: CellFlowWidget.cpp:415-421.  `2.0f * M_PI * lfoS * t` is evaluated in double (M_PI),
 * sin() in double, the product with lfoA in double, then rounded into the float field. */
float orc_ratio_with_lfo(const cf_params* p, float t) {
    if (p->lfoA != 0.0f) {
        double lfo = (double)p->lfoA * sin(2.0f * M_PI * (double)p->lfoS * (double)t);
        return (float)((double)p->ratio + lfo);
    }
    return p->ratio;
}

/* ------------------------------------------------------------------------------------------
 * CUDA expf as emitted for .cu:120 (PTX: fma.rn / cvt.sat / fma.rm / ex2.approx.ftz):
 *   a = sat(fma(t, 0x3BBB989D, 0.5));  b = fma_rd(a, 252, 12582913);  c = b - 12583039;
 *   q = fma(t, 1.4426950216, -c);  q = fma(t, 1.9259630335e-8, q);
 *   e = ex2(q) * float_from_bits(bits(b) << 23)
 * ---------------------------------------------------------------------------------------- */
static inline float bits_f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t f_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float fma_rd(float a, float b, float c) {
    /* a in [0,1] (24 bits) times 252 plus 12582913 is exact in double; round toward -inf. */
    double d = (double)a * (double)b + (double)c;
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
}
static inline float cuda_expf(float t) {
    float a = fmaf(t, bits_f(0x3BBB989Du), 0.5f);
    a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);
    if (a != a) a = 0.0f;
    float b = fma_rd(a, 252.0f, 12582913.0f);
    float c = b - 12583039.0f;
    float q = fmaf(t, bits_f(0x3FB8AA3Bu), -c);
    q = fmaf(t, bits_f(0x32A57060u), q);
    float scale = bits_f(f_bits(b) << 23);
    return exp2f(q) * scale;
}

/* Effective radius of a type pair, .cu:108-110 as compiled:
 *   a_x = fma(radio[x], ratio, 1);  Reff = fma(a_p, radius, a_o * radius) * 0.5 */
static inline float reff_pair(const cf_params* p, const float* radio, uint32_t tp, uint32_t to) {
    float ap = fmaf(radio[tp], p->ratioWithLFO, 1.0f);
    float ao = fmaf(radio[to], p->ratioWithLFO, 1.0f);
    return fmaf(ap, p->radius, ao * p->radius) * 0.5f;
}
void orc_reff_table(const cf_params* p, const float* radio, float* out) {
    int T = p->numParticleTypes;
    for (int a = 0; a < T; a++)
        for (int b = 0; b < T; b++) out[a * T + b] = reff_pair(p, radio, a, b);
}

/* Minimum-image displacement, .cu:92-102 (second test sees the result of the first). */
static inline float wrap_delta(float d, float W) {
    if (d > W * 0.5f) d = d - W;
    if (d < W * -0.5f) d = d + W;
    return d;
}

typedef struct pair_acc {
    float fx, fy, fz;
    float fabs_sum; /* sum over accepted pairs of |forceValue| * (|repulsion term| + |attraction
                       term|): the magnitude force errors are measured against.  The net of the
                       two terms is NOT a usable scale: it crosses zero inside the radius, and
                       there even the reference's own fp32 arithmetic is off by more than 1e-5 of
                       it (tests/test_oracle.py::test_f64_error_budget). */
    float fnet_sum; /* sum over accepted pairs of |forceValue * netForce| = sum |f_ij| (SURVEY.md section 7's norm);
                       reported beside the gross-term norm, not used as the tolerance scale (see above) */
    int count;
} pair_acc;

/* One ordered pair (p <- o), .cu:92-131.  Returns 1 if accepted. */
static inline int pair_term(const cf_params* P, const float* table, const float* reffT,
                            const cf_particle* p, const cf_particle* o, pair_acc* acc) {
    int T = P->numParticleTypes;
    float dx = wrap_delta(o->pos[0] - p->pos[0], P->canvasWidth);
    float dy = wrap_delta(o->pos[1] - p->pos[1], P->canvasHeight);
    float dz = wrap_delta(o->pos[2] - p->pos[2], P->canvasDepth);
    float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
    float dist = sqrtf(d2 + 0.0001f);
    float reff = reffT[p->ptype * T + o->ptype];
    if (!(dist < reff)) return 0;
    acc->count++;
    float fv = table[p->ptype * T + o->ptype];
    float r = dist / reff;
    float t = (r * -P->k) * r;
    float e = cuda_expf(t);
    float net = fmaf(e, P->repulsion, -(r * P->attraction));
    float s = fv * net;
    acc->fx = fmaf(s, dx / dist, acc->fx);
    acc->fy = fmaf(s, dy / dist, acc->fy);
    acc->fz = fmaf(s, dz / dist, acc->fz);
    acc->fabs_sum += fabsf(fv) * (fabsf(e * P->repulsion) + fabsf(r * P->attraction));
    acc->fnet_sum += fabsf(s);
    return 1;
}

/* Per-particle epilogue, .cu:136-165 as compiled. */
static inline void finish_particle(const cf_params* P, const cf_particle* p, int prev,
                                   const pair_acc* acc, cf_particle* out) {
    float avg = (float)(acc->count + prev) * 0.5f;
    float dens = fminf(avg / (float)P->maxExpectedNeighbors, 1.0f);
    float a = fmaf(dens, -(1.0f - P->balance), 1.0f);
    float m = a * P->forceMultiplier;
    float f[3] = {m * acc->fx, m * acc->fy, m * acc->fz};
    float W[3] = {P->canvasWidth, P->canvasHeight, P->canvasDepth};
    *out = *p;
    for (int c = 0; c < 3; c++) {
        out->acc[c] = f[c];
        float v = fmaf(p->vel[c], P->friction, f[c] * P->delta_t);
        out->vel[c] = v;
        float x = fmaf(v, P->delta_t, p->pos[c]);
        out->pos[c] = fmodf(x + W[c], W[c]);
    }
}

/* Brute-force step exactly as the reference loops (.cu:86: j = 0..N-1, j != i). */
void orc_step_bruteforce(const cf_particle* in, const int32_t* cnt_in, int n, const cf_params* P,
                         const float* table, const float* radio, cf_particle* out,
                         int32_t* cnt_out, float* fabs_out, int nthreads) {
    float reffT[CF_MAX_PARTICLE_TYPES * CF_MAX_PARTICLE_TYPES];
    orc_reff_table(P, radio, reffT);
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int i = 0; i < n; i++) {
        pair_acc acc = {0.f, 0.f, 0.f, 0.f, 0.f, 0};
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            pair_term(P, table, reffT, &in[i], &in[j], &acc);
        }
        finish_particle(P, &in[i], cnt_in ? cnt_in[i] : 0, &acc, &out[i]);
        cnt_out[i] = acc.count;
        if (fabs_out) fabs_out[i] = acc.fabs_sum;
    }
}

/* ------------------------------------------------------------------------------------------
 * Cell-accelerated step with results IDENTICAL to orc_step_bruteforce: candidates come from a
 * uniform grid, are sorted by particle index, and then go through the same pair_term in the
 * same j order.  Only a test-speed device; the grid here is the oracle's own (edge >= R_max).
 * ---------------------------------------------------------------------------------------- */
typedef struct grid {
    int nx, ny, nz;
    int* start; /* ncell+1 */
    int* items;
} grid;

static float max_reff(const cf_params* P, const float* radio) {
    float reffT[CF_MAX_PARTICLE_TYPES * CF_MAX_PARTICLE_TYPES];
    orc_reff_table(P, radio, reffT);
    float m = 0.f;
    int T = P->numParticleTypes;
    for (int i = 0; i < T * T; i++)
        if (reffT[i] > m) m = reffT[i];
    return m;
}

static int cell_of(float x, float W, int n) {
    int c = (int)floor((double)x / (double)W * n);
    if (c < 0) c = 0;
    if (c >= n) c = n - 1;
    return c;
}

static void grid_build(grid* g, const cf_particle* in, int n, const float W[3], float edge_min) {
    int dims[3];
    for (int c = 0; c < 3; c++) {
        /* a margin keeps every in-range partner inside the 27 cells whatever the rounding */
        int k = (int)floor((double)W[c] / ((double)edge_min * 1.001 + 1e-3));
        if (k < 1) k = 1;
        if (k > 256) k = 256;
        dims[c] = k;
    }
    g->nx = dims[0];
    g->ny = dims[1];
    g->nz = dims[2];
    int nc = g->nx * g->ny * g->nz;
    g->start = (int*)calloc((size_t)nc + 1, sizeof(int));
    g->items = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    int* cell = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; i++) {
        int cx = cell_of(in[i].pos[0], W[0], g->nx), cy = cell_of(in[i].pos[1], W[1], g->ny),
            cz = cell_of(in[i].pos[2], W[2], g->nz);
        cell[i] = (cx * g->ny + cy) * g->nz + cz;
        g->start[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; c++) g->start[c + 1] += g->start[c];
    int* fill = (int*)malloc((size_t)nc * sizeof(int));
    memcpy(fill, g->start, (size_t)nc * sizeof(int));
    for (int i = 0; i < n; i++) g->items[fill[cell[i]]++] = i; /* ascending index per cell */
    free(fill);
    free(cell);
}
static void grid_free(grid* g) {
    free(g->start);
    free(g->items);
}

/* Candidate lists are sorted by particle index so that the cell path sums in the brute-force
 * order (bit-identical results).  Large tolerance-based runs may switch the sort off. */
static int g_sort_candidates = 1;
void orc_set_sort_candidates(int on) { g_sort_candidates = on; }

static int cmp_int(const void* a, const void* b) {
    int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}

/* Distinct neighbour cell coordinates along one axis (periodic or clamped). */
static int axis_neighbours(int c, int n, int periodic, int out[3]) {
    int k = 0;
    for (int d = -1; d <= 1; d++) {
        int v = c + d;
        if (periodic) {
            v = ((v % n) + n) % n;
        } else if (v < 0 || v >= n) {
            continue;
        }
        int dup = 0;
        for (int q = 0; q < k; q++)
            if (out[q] == v) dup = 1;
        if (!dup) out[k++] = v;
    }
    return k;
}

/* Gather candidate indices for particle i (sorted ascending). Returns count. */
static int gather_candidates(const grid* g, const cf_particle* in, int i, const float W[3],
                             int periodic, int** buf, int* cap) {
    int cx = cell_of(in[i].pos[0], W[0], g->nx), cy = cell_of(in[i].pos[1], W[1], g->ny),
        cz = cell_of(in[i].pos[2], W[2], g->nz);
    int ax[3], ay[3], az[3];
    int kx = axis_neighbours(cx, g->nx, periodic, ax), ky = axis_neighbours(cy, g->ny, periodic, ay),
        kz = axis_neighbours(cz, g->nz, periodic, az);
    int m = 0;
    for (int a = 0; a < kx; a++)
        for (int b = 0; b < ky; b++)
            for (int c = 0; c < kz; c++) {
                int cell = (ax[a] * g->ny + ay[b]) * g->nz + az[c];
                int s = g->start[cell], e = g->start[cell + 1];
                if (m + (e - s) > *cap) {
                    *cap = (m + (e - s)) * 2 + 64;
                    *buf = (int*)realloc(*buf, (size_t)*cap * sizeof(int));
                }
                memcpy(*buf + m, g->items + s, (size_t)(e - s) * sizeof(int));
                m += e - s;
            }
    if (g_sort_candidates) qsort(*buf, (size_t)m, sizeof(int), cmp_int);
    return m;
}

void orc_step_cells_range(const cf_particle* in, const int32_t* cnt_in, int n, int i0, int i1,
                          const cf_params* P, const float* table, const float* radio,
                          cf_particle* out, int32_t* cnt_out, float* fabs_out, int nthreads);

/* Optional second norm of the next cell-list steps: fnet[i] = sum_j |f_ij| (NULL switches it off). */
static float* g_fnet_out = NULL;
void orc_set_fnet_output(float* fnet) { g_fnet_out = fnet; }

void orc_step_cells(const cf_particle* in, const int32_t* cnt_in, int n, const cf_params* P,
                    const float* table, const float* radio, cf_particle* out, int32_t* cnt_out,
                    float* fabs_out, int nthreads) {
    orc_step_cells_range(in, cnt_in, n, 0, n, P, table, radio, out, cnt_out, fabs_out, nthreads);
}

/* Particles [i0, i1) only (all n particles act on them): the bounded sample bench.py times. */
void orc_step_cells_range(const cf_particle* in, const int32_t* cnt_in, int n, int i0, int i1,
                          const cf_params* P, const float* table, const float* radio,
                          cf_particle* out, int32_t* cnt_out, float* fabs_out, int nthreads) {
    float reffT[CF_MAX_PARTICLE_TYPES * CF_MAX_PARTICLE_TYPES];
    orc_reff_table(P, radio, reffT);
    float W[3] = {P->canvasWidth, P->canvasHeight, P->canvasDepth};
    float rmax = max_reff(P, radio);
    if (!(rmax > 0.f)) rmax = 1.0f;
    grid g;
    grid_build(&g, in, n, W, rmax);
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        int cap = 1024;
        int* buf = (int*)malloc((size_t)cap * sizeof(int));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int i = i0; i < i1; i++) {
            pair_acc acc = {0.f, 0.f, 0.f, 0.f, 0.f, 0};
            int m = gather_candidates(&g, in, i, W, 1, &buf, &cap);
            for (int q = 0; q < m; q++) {
                int j = buf[q];
                if (j == i) continue;
                pair_term(P, table, reffT, &in[i], &in[j], &acc);
            }
            finish_particle(P, &in[i], cnt_in ? cnt_in[i] : 0, &acc, &out[i]);
            cnt_out[i] = acc.count;
            if (fabs_out) fabs_out[i] = acc.fabs_sum;
            if (g_fnet_out) g_fnet_out[i] = acc.fnet_sum;
        }
        free(buf);
    }
    grid_free(&g);
}

/* ------------------------------------------------------------------------------------------
 * fp64 restatement of the same law (error budgeting only): real-number formulas of
 * .cu:92-161 in double with libm exp; the neighbour DECISION is taken from the fp32 law above
 * so that both sum over the same pair set.
 * out_force[3n], out_pos[3n] (position before the fmod wrap is applied in double).
 * ---------------------------------------------------------------------------------------- */
void orc_step_f64(const cf_particle* in, const int32_t* cnt_in, int n, const cf_params* P,
                  const float* table, const float* radio, double* out_force, double* out_pos,
                  double* out_fabs, int nthreads) {
    float reffT[CF_MAX_PARTICLE_TYPES * CF_MAX_PARTICLE_TYPES];
    orc_reff_table(P, radio, reffT);
    float W[3] = {P->canvasWidth, P->canvasHeight, P->canvasDepth};
    int T = P->numParticleTypes;
    float rmax = max_reff(P, radio);
    if (!(rmax > 0.f)) rmax = 1.0f;
    grid g;
    grid_build(&g, in, n, W, rmax);
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        int cap = 1024;
        int* buf = (int*)malloc((size_t)cap * sizeof(int));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int i = 0; i < n; i++) {
            const cf_particle* p = &in[i];
            double F[3] = {0, 0, 0}, fabs_sum = 0;
            int count = 0;
            int m = gather_candidates(&g, in, i, W, 1, &buf, &cap);
            for (int q = 0; q < m; q++) {
                int j = buf[q];
                if (j == i) continue;
                const cf_particle* o = &in[j];
                float dxf = wrap_delta(o->pos[0] - p->pos[0], W[0]);
                float dyf = wrap_delta(o->pos[1] - p->pos[1], W[1]);
                float dzf = wrap_delta(o->pos[2] - p->pos[2], W[2]);
                float d2f = fmaf(dzf, dzf, fmaf(dxf, dxf, dyf * dyf));
                float reff = reffT[p->ptype * T + o->ptype];
                if (!(sqrtf(d2f + 0.0001f) < reff)) continue;
                count++;
                double d[3];
                for (int c = 0; c < 3; c++) {
                    double dd = (double)o->pos[c] - (double)p->pos[c];
                    if (dd > 0.5 * W[c]) dd -= W[c];
                    if (dd < -0.5 * W[c]) dd += W[c];
                    d[c] = dd;
                }
                double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + (double)0.0001f);
                double r = dist / (double)reff;
                double net = (double)P->repulsion * exp(-(double)P->k * r * r) -
                             (double)P->attraction * r;
                double s = net * (double)table[p->ptype * T + o->ptype];
                for (int c = 0; c < 3; c++) F[c] += d[c] / dist * s;
                fabs_sum += fabs((double)table[p->ptype * T + o->ptype]) *
                            (fabs((double)P->repulsion * exp(-(double)P->k * r * r)) +
                             fabs((double)P->attraction * r));
            }
            int prev = cnt_in ? cnt_in[i] : 0;
            double avg = (double)(count + prev) * 0.5;
            double dens = fmin(avg / (double)P->maxExpectedNeighbors, 1.0);
            double a = 1.0 - (1.0 - (double)P->balance) * dens;
            double mlt = (double)P->forceMultiplier * a;
            for (int c = 0; c < 3; c++) {
                double f = F[c] * mlt;
                double v = (double)p->vel[c] * (double)P->friction + f * (double)P->delta_t;
                out_force[3 * i + c] = f;
                out_pos[3 * i + c] = (double)p->pos[c] + v * (double)P->delta_t;
            }
            if (out_fabs) out_fabs[i] = fabs_sum * fabs(mlt);
        }
        free(buf);
    }
    grid_free(&g);
}

/* ------------------------------------------------------------------------------------------
 * Proximity graph, .cu:212-246: for particle i scan j = i+1.. in index order while fewer than
 * 2*maxConn hits; same type only; plain (non-wrapped) distance, d2 = fma(dz,dz,fma(dx,dx,dy*dy))
 * as compiled; strict d2 < dist^2; stable insertion sort by d2; keep the first maxConn.
 * Emits edges (i, j) in particle order.  Returns the edge count.
 * ---------------------------------------------------------------------------------------- */
static int graph_select(const cf_particle* in, int i, const int* cand, int m, float dist2,
                        int maxConn, cf_edge* out) {
    float nd[2 * CF_MAX_GRAPH_CONN];
    int ni[2 * CF_MAX_GRAPH_CONN];
    int nearby = 0;
    const cf_particle* p1 = &in[i];
    for (int q = 0; q < m && nearby < maxConn * 2; q++) {
        int j = cand ? cand[q] : i + 1 + q;
        if (j <= i) continue;
        const cf_particle* p2 = &in[j];
        if (p1->ptype != p2->ptype) continue;
        float dx = p2->pos[0] - p1->pos[0];
        float dy = p2->pos[1] - p1->pos[1];
        float dz = p2->pos[2] - p1->pos[2];
        float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        if (d2 < dist2) {
            nd[nearby] = d2;
            ni[nearby] = j;
            nearby++;
        }
    }
    for (int a = 1; a < nearby; a++) { /* .cu:235-243 */
        float kd = nd[a];
        int ki = ni[a];
        int b = a - 1;
        while (b >= 0 && nd[b] > kd) {
            nd[b + 1] = nd[b];
            ni[b + 1] = ni[b];
            b--;
        }
        nd[b + 1] = kd;
        ni[b + 1] = ki;
    }
    int w = nearby < maxConn ? nearby : maxConn;
    for (int a = 0; a < w; a++) {
        out[a].i = i;
        out[a].j = ni[a];
    }
    return w;
}

int orc_graph_bruteforce(const cf_particle* in, int n, float proximity_distance, int max_conn,
                         cf_edge* edges, int capacity) {
    int maxConn = max_conn < CF_MAX_GRAPH_CONN ? max_conn : CF_MAX_GRAPH_CONN;
    float dist2 = proximity_distance * proximity_distance; /* .cu:660 */
    int ne = 0;
    cf_edge tmp[CF_MAX_GRAPH_CONN];
    for (int i = 0; i < n; i++) {
        int w = graph_select(in, i, NULL, n - i - 1, dist2, maxConn, tmp);
        for (int a = 0; a < w; a++) {
            if (ne < capacity) edges[ne] = tmp[a];
            ne++;
        }
    }
    return ne;
}

/* Same edge list through a (non-periodic) grid; identical output to the brute-force form. */
int orc_graph_cells(const cf_particle* in, int n, const float canvas[3], float proximity_distance,
                    int max_conn, cf_edge* edges, int capacity) {
    int maxConn = max_conn < CF_MAX_GRAPH_CONN ? max_conn : CF_MAX_GRAPH_CONN;
    float dist2 = proximity_distance * proximity_distance;
    float edge = proximity_distance > 1.0f ? proximity_distance : 1.0f;
    grid g;
    grid_build(&g, in, n, canvas, edge);
    int ne = 0;
    int cap = 1024;
    int* buf = (int*)malloc((size_t)cap * sizeof(int));
    cf_edge tmp[CF_MAX_GRAPH_CONN];
    for (int i = 0; i < n; i++) {
        int m = gather_candidates(&g, in, i, canvas, 0, &buf, &cap);
        int w = graph_select(in, i, buf, m, dist2, maxConn, tmp);
        for (int a = 0; a < w; a++) {
            if (ne < capacity) edges[ne] = tmp[a];
            ne++;
        }
    }
    free(buf);
    grid_free(&g);
    return ne;
}

/* Reference VBO layout of one edge list, .cu:255-275: 12 floats per edge. */
void orc_graph_vertices(const cf_particle* in, const cf_edge* edges, int ne, const cf_color* colors,
                        int num_types, float* out) {
    for (int e = 0; e < ne; e++) {
        const cf_particle* a = &in[edges[e].i];
        const cf_particle* b = &in[edges[e].j];
        const cf_color* c = &colors[a->ptype % (uint32_t)num_types]; /* .cu:253 */
        float* v = out + 12 * (size_t)e;
        v[0] = a->pos[0], v[1] = a->pos[1], v[2] = a->pos[2];
        v[3] = c->r, v[4] = c->g, v[5] = c->b;
        v[6] = b->pos[0], v[7] = b->pos[1], v[8] = b->pos[2];
        v[9] = c->r, v[10] = c->g, v[11] = c->b;
    }
}

/* moveParticlesKernel, .cu:182-184: pos = fmodf(pos + d + W, W), evaluated left to right. */
void orc_move_universe(cf_particle* p, int n, float dx, float dy, float dz, const float W[3]) {
    float d[3] = {dx, dy, dz};
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) p[i].pos[c] = fmodf((p[i].pos[c] + d[c]) + W[c], W[c]);
}

/* ------------------------------------------------------------------------------------------
 * Initial conditions.  The reference seeds cuRAND with time(nullptr) (.cu:480), so no input of
 * its own is reproducible; the engine replaces that by a counter-based generator keyed by
 * (seed, particle id) with the reference's spawn SHAPE (.cu:45-62).  This restates the
 * engine's generator (cellflow_b200/csrc/init.cuh) so tests can check it bit-for-bit.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float u01(uint64_t h) { return (float)(h >> 40) * 5.9604644775390625e-08f; }

void orc_init_particles(cf_particle* out, int n, int id0, int T, uint64_t seed, int mode,
                        const float W[3]) {
    for (int k = 0; k < n; k++) {
        uint64_t id = (uint64_t)(id0 + k);
        uint64_t s = mix64(seed ^ mix64(id));
        cf_particle* p = &out[k];
        memset(p, 0, sizeof(*p));
        for (int c = 0; c < 3; c++) {
            float span = mode == CF_INIT_SPAWN_CUBE ? fminf(2000.0f, W[c]) : W[c];
            float off = (W[c] - span) * 0.5f;
            float x = fmaf(u01(mix64(s + (uint64_t)c)), span, off);
            if (x >= W[c]) x = nextafterf(W[c], 0.0f);
            p->pos[c] = x;
        }
        uint32_t t = (uint32_t)(u01(mix64(s + 3u)) * (float)T);
        p->ptype = t < (uint32_t)T ? t : (uint32_t)T - 1u;
    }
}

/* ------------------------------------------------------------------------------------------
 * Cell assignment of the engine's grid (new; no reference counterpart — SURVEY.md row N1).
 * Restates cellflow_b200/csrc/grid.cuh: cell coordinate = min((int)(pos * inv), n-1) with
 * inv = (float)n / W, linear key = (cx*ny + cy)*nz + cz.
 * ---------------------------------------------------------------------------------------- */
void orc_cell_keys(const cf_particle* in, int n, const float W[3], const int dims[3],
                   uint32_t* keys) {
    float inv[3];
    for (int c = 0; c < 3; c++) inv[c] = (float)dims[c] / W[c];
    for (int i = 0; i < n; i++) {
        int cc[3];
        for (int c = 0; c < 3; c++) {
            int v = (int)(in[i].pos[c] * inv[c]);
            if (v > dims[c] - 1) v = dims[c] - 1;
            if (v < 0) v = 0;
            cc[c] = v;
        }
        keys[i] = (uint32_t)((cc[0] * dims[1] + cc[1]) * dims[2] + cc[2]);
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
