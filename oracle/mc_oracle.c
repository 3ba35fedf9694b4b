/* ORACLE (test infrastructure, not product code): scalar CPU restatement of the reference's sparse
 * marching cubes with cross-PLIVox std-weighted blending.
 *
 *   reference: pytorch/system/ext/marching_cubes/mc_interp_kernel.cu
 *       query_sdf_raw  :7-29     -> raw_lookup()
 *       get_sdf        :34-185   -> blended_corner()
 *       sdf_interp     :187-200  -> edge_vertex()
 *       meshing_cube   :202-320  -> dif_oracle_marching_cubes() body
 *       host wrapper   :322-382  -> count semantics (triangles past max_tri are counted, not written)
 *   tables: include/dif_mc_tables.h (values == mc_data.cuh:40,54)
 *
 * Parity status: PINNED against executions of the unmodified reference kernel.  oracle/build_ref.py compiles
 * the reference extension from /root/reference into oracle/_ref/marching_cubes/, tests/golden/make_golden_gpu.py
 * runs it on a B200 and stores inputs + outputs in tests/golden/ref_ext_mc_r*.npz; tests/test_oracle_mc.py checks
 * this restatement against those vectors BIT-EXACTLY, and tests/test_ref_ext_gpu.py compares the product kernel with
 * the reference module live on the GPU.
 *
 * Rounding: built with -ffp-contract=off; every float op is a separately rounded IEEE fp32 op EXCEPT where nvcc
 * (default -fmad=true) contracts the reference source into a fused multiply-add.  Those places were read off the
 * PTX/SASS nvcc 12.9 emits for the reference file and are written as explicit fmaf() here (and as fmaf in the
 * product kernel):   total_sdf.x += (sdf*w)*std  -> fma(sdf*w, std, acc)      (:109, :118, ... one per corner)
 *                    total_weight.x / total_sdf.y += w*std -> fma(w, std, acc) (one register: the compiler merges both)
 *                    total_weight.y += w        -> plain add
 *                    points[i] = bpos + rpos*sbs -> fma(rpos, sbs, bpos)       (:224-252)
 *                    p1*w1 + p2*w2              -> fma(p2, w2, p1*w1)          (:195-198)
 * Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may load this.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/dif_mc_tables.h"

typedef struct { float sdf, std; } sv_t;

typedef struct {
    const int64_t* indexer; int nx, ny, nz;
    const int32_t* mapping; int64_t mapping_len;
    const float* cube_sdf; const float* cube_std; int r;   /* cubes are [B][2r][2r][2r] */
} mc_ctx;

/* mc_interp_kernel.cu:7-29.  Coordinates are unsigned in the reference, so "-1" wraps and fails the range test. */
static sv_t raw_lookup(const mc_ctx* c, uint32_t bx, uint32_t by, uint32_t bz, uint32_t ax, uint32_t ay, uint32_t az) {
    sv_t miss = { NAN, NAN };
    if (bx >= (uint32_t)c->nx || by >= (uint32_t)c->ny || bz >= (uint32_t)c->nz) return miss;
    int64_t slot = c->indexer[((int64_t)bx * c->ny + by) * c->nz + bz];
    if (slot == -1 || slot >= c->mapping_len) return miss;
    int32_t b = c->mapping[slot];
    if (b == -1) return miss;
    const int n = 2 * c->r;
    const int64_t off = (((int64_t)b * n + ax) * n + ay) * n + az;
    sv_t v = { c->cube_sdf[off], c->cube_std[off] };
    return v;
}

/* mc_interp_kernel.cu:34-185 */
static sv_t blended_corner(const mc_ctx* c, uint32_t bx, uint32_t by, uint32_t bz, uint32_t px, uint32_t py, uint32_t pz) {
    const uint32_t r = (uint32_t)c->r;
    uint32_t bpos[3] = { bx, by, bz }, rpos[3] = { px, py, pz };
    const uint32_t bsize[3] = { (uint32_t)c->nx, (uint32_t)c->ny, (uint32_t)c->nz };
    for (int a = 0; a < 3; ++a) if (bpos[a] >= bsize[a]) { bpos[a] = bsize[a] - 1; rpos[a] = r - 1; }

    const uint32_t rbound = (r - 1) / 2, rstart = r / 2;
    const float rmid = r / 2.0f;
    float w_m[3], w_p[3]; int b_m[3], b_p[3], r_m[3], r_p[3], own_is_p[3];
    for (int a = 0; a < 3; ++a) {
        if (rpos[a] <= rbound) {
            b_m[a] = -1; r_m[a] = (int)r; b_p[a] = 0; r_p[a] = 0;
            w_p[a] = (float)rpos[a] + rmid; w_m[a] = rmid - (float)rpos[a];
            own_is_p[a] = 1;
        } else {
            b_m[a] = 0; r_m[a] = 0; b_p[a] = 1; r_p[a] = -(int)r;
            w_p[a] = (float)rpos[a] - rmid; w_m[a] = rmid + (float)r - (float)rpos[a];
            own_is_p[a] = 0;
        }
        w_m[a] /= (float)r; w_p[a] /= (float)r;
        rpos[a] += rstart;
    }
    const int own = own_is_p[0] * 4 + own_is_p[1] * 2 + own_is_p[2];

    float s1 = 0.f, s2 = 0.f, s4 = 0.f;
    for (int k = 0; k < 8; ++k) {                         /* order mmm, mmp, mpm, mpp, pmm, pmp, ppm, ppp */
        const int xp = (k >> 2) & 1, yp = (k >> 1) & 1, zp = k & 1;
        sv_t q = raw_lookup(c,
            bpos[0] + (uint32_t)(xp ? b_p[0] : b_m[0]), bpos[1] + (uint32_t)(yp ? b_p[1] : b_m[1]), bpos[2] + (uint32_t)(zp ? b_p[2] : b_m[2]),
            rpos[0] + (uint32_t)(xp ? r_p[0] : r_m[0]), rpos[1] + (uint32_t)(yp ? r_p[1] : r_m[1]), rpos[2] + (uint32_t)(zp ? r_p[2] : r_m[2]));
        float w = (xp ? w_p[0] : w_m[0]) * (yp ? w_p[1] : w_m[1]);
        w = w * (zp ? w_p[2] : w_m[2]);
        if (!isnan(q.sdf)) {
            const float t = q.sdf * w;
            s1 = fmaf(t, q.std, s1);                       /* total_sdf.x */
            s2 = fmaf(w, q.std, s2);                       /* total_weight.x == total_sdf.y */
            s4 += w;                                       /* total_weight.y */
        } else if (own == k) {
            sv_t miss = { NAN, NAN };
            return miss;
        }
    }
    sv_t out = { s1 / s2, s2 / s4 };
    return out;
}

typedef struct { float x, y, z, w; } v4;

/* mc_interp_kernel.cu:187-200 */
static v4 edge_vertex(const float* p1, const float* p2, float std1, float std2, float v1, float v2) {
    v4 a = { p1[0], p1[1], p1[2], std1 }, b = { p2[0], p2[1], p2[2], std2 };
    if (fabsf(0.0f - v1) < 1.0e-5f) return a;
    if (fabsf(0.0f - v2) < 1.0e-5f) return b;
    if (fabsf(v1 - v2) < 1.0e-5f) return a;
    float w2 = (0.0f - v1) / (v2 - v1);
    float w1 = 1 - w2;
    v4 o;
    o.x = fmaf(p2[0], w2, p1[0] * w1);
    o.y = fmaf(p2[1], w2, p1[1] * w1);
    o.z = fmaf(p2[2], w2, p1[2] * w1);
    o.w = fmaf(std2, w2, std1 * w1);
    return o;
}

static const int CORNER[8][3] = { {0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1} };   /* :240-270 */
static const int EDGE[12][2] = { {0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7} }; /* :284-295 */

/* Returns the number of triangles produced (may exceed max_tri; only the first max_tri are written). */
int64_t dif_oracle_marching_cubes(const int64_t* indexer, const int64_t* valid_blocks, int64_t n_blocks,
                                  const int32_t* mapping, int64_t mapping_len,
                                  const float* cube_sdf, const float* cube_std, int r,
                                  int nx, int ny, int nz, float max_std, int64_t max_tri,
                                  float* tri /*[max_tri][3][3]*/, int64_t* tri_id /*[max_tri]*/, float* tri_std /*[max_tri][3]*/) {
    mc_ctx c = { indexer, nx, ny, nz, mapping, mapping_len, cube_sdf, cube_std, r };
    const float sbs = 1.0f / (float)r;
    int64_t count = 0;
    for (int64_t k = 0; k < n_blocks; ++k) {
        const int64_t id = valid_blocks[k];
        const uint32_t bx = (uint32_t)((id / ((int64_t)ny * nz)) % nx), by = (uint32_t)((id / nz) % ny), bz = (uint32_t)(id % nz);
        for (uint32_t sub = 0; sub < (uint32_t)(r * r * r); ++sub) {
            const uint32_t rx = sub / (r * r), ry = (sub / r) % r, rz = sub % r;
            sv_t val[8]; float pt[8][3]; int bad = 0;
            for (int i = 0; i < 8 && !bad; ++i) {
                const uint32_t cx = rx + CORNER[i][0], cy = ry + CORNER[i][1], cz = rz + CORNER[i][2];
                val[i] = blended_corner(&c, bx, by, bz, cx, cy, cz);
                if (isnan(val[i].sdf)) { bad = 1; break; }
                pt[i][0] = fmaf((float)cx, sbs, (float)bx); pt[i][1] = fmaf((float)cy, sbs, (float)by); pt[i][2] = fmaf((float)cz, sbs, (float)bz);
            }
            if (bad) continue;
            int type = 0;
            for (int i = 0; i < 8; ++i) if (val[i].sdf < 0) type |= 1 << i;
            const int emask = dif_mc_edge_mask[type];
            if (emask == 0) continue;
            v4 vert[12];
            for (int e = 0; e < 12; ++e) if (emask & (1 << e)) {
                const int a = EDGE[e][0], b = EDGE[e][1];
                vert[e] = edge_vertex(pt[a], pt[b], val[a].std, val[b].std, val[a].sdf, val[b].sdf);
            }
            for (int i = 0; dif_mc_tri_edges[type][i] != -1; i += 3) {
                v4 vp[3];
                for (int vi = 0; vi < 3; ++vi) vp[vi] = vert[dif_mc_tri_edges[type][i + vi]];
                if (vp[0].w > max_std || vp[1].w > max_std || vp[2].w > max_std) continue;
                const int64_t t = count++;
                if (t < max_tri) {
                    for (int vi = 0; vi < 3; ++vi) {
                        tri[(t * 3 + vi) * 3 + 0] = vp[vi].x; tri[(t * 3 + vi) * 3 + 1] = vp[vi].y; tri[(t * 3 + vi) * 3 + 2] = vp[vi].z;
                        tri_std[t * 3 + vi] = vp[vi].w;
                    }
                    tri_id[t] = id;
                }
            }
        }
    }
    return count;
}
