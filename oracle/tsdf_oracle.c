/*
 * oracle/tsdf_oracle.c -- CPU restatement of the reference TSDF integration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under union_thesis_slam_b200/ may import,
 * link or call this file; it is the checker for tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * sha256 digests and counts produced by the unmodified reference
 * (/root/reference/grid_fusion.py, hash_fusion.py run in the build container by
 * tests/golden/make_golden.py) on the bundled RGB-D frames and on small random
 * volumes, and against the call counts frozen in the reference's own committed
 * profiler dumps (354 383 / 389 590 / 3 557 017).
 *
 * What it restates (reference file:line, /root/reference):
 *   grid_fusion.py:170-181  vox2world      f32( f64(origin) + vs*f64(f32(idx)) )
 *   grid_fusion.py:363-368  rigid_transform  np.dot(Tinv, [x y z 1]) == dgemm:
 *                           acc = t0*x ; acc = fma(t1,y,acc) ; fma(t2,z,acc) ; fma(t3,1,acc)
 *   grid_fusion.py:183-197  cam2pix        rint((X*fx)/Z + cx), f64, ties-to-even
 *   grid_fusion.py:273-286  frustum/depth masks, dist = min(1, diff/trunc)
 *   grid_fusion.py:199-212  integrate_tsdf (f32 product, f64 sum/divide, f32 store)
 *   grid_fusion.py:302-314  colour blend, all f32, np.round == rintf
 *   grid_fusion.py:228-232  colour fold
 *   hash_fusion.py:103-145  HashTable.integrate: same pre-pass, then per-voxel
 *   data_structures/voxel.py:19-49  Voxel.integrate (float64 running mean, obs_weight
 *                           fixed at its default 1.0 because hash_fusion.py:141,145
 *                           never forwards obs_weight)
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -march=x86-64-v3 -shared -fPIC  (oracle/Makefile)
 * -ffp-contract=off is REQUIRED: every contraction below is an explicit fma().
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* grid_fusion.py:228-232 -- floor(B*65536 + G*256 + R) in f32 (exact integers). */
void oracle_fold_color(const uint8_t *rgb, float *out, int64_t npix)
{
    for (int64_t p = 0; p < npix; ++p) {
        float r = (float)rgb[3 * p + 0], g = (float)rgb[3 * p + 1], b = (float)rgb[3 * p + 2];
        out[p] = floorf(b * 65536.0f + g * 256.0f + r);
    }
}

/* grid_fusion.py:175-180: world coordinate of lattice index along one axis. */
static inline double lattice_coord(float origin, double vs, int64_t idx)
{
    float w = (float)((double)origin + vs * (double)(float)idx);
    return (double)w;
}

typedef struct {
    double fx, fy, cx, cy; /* f64(f32(K)) -- grid_fusion.py:190-192 */
    double T[16];          /* np.linalg.inv(cam_pose), row major, computed by the caller */
    double trunc;
    int H, W;
    const double *depth;   /* metres, 0 = invalid */
    const float *cfold;    /* folded colour image */
} frame_t;

/* Steps 3-7 of SURVEY Appendix A.  Returns 1 and fills dist/pix when the voxel is updated. */
static inline int project_voxel(const frame_t *f, double cxr[3], double wz,
                                double *dist, int64_t *pix)
{
    /* cxr[r] = fma(T[r][1], wy, T[r][0]*wx) is the per-row partial of the dgemm chain */
    double X = fma(f->T[3], 1.0, fma(f->T[2], wz, cxr[0]));
    double Y = fma(f->T[7], 1.0, fma(f->T[6], wz, cxr[1]));
    double Z = fma(f->T[11], 1.0, fma(f->T[10], wz, cxr[2]));
    if (!(Z > 0.0)) return 0;                                /* grid_fusion.py:277 */
    double u = rint((X * f->fx) / Z + f->cx);                /* :195 */
    double v = rint((Y * f->fy) / Z + f->cy);                /* :196 */
    if (!(u >= 0.0 && u < (double)f->W && v >= 0.0 && v < (double)f->H)) return 0; /* :273-276 */
    int64_t p = (int64_t)v * f->W + (int64_t)u;
    double d = f->depth[p];
    if (!(d > 0.0)) return 0;                                /* :284 depth_val > 0 */
    double diff = d - Z;                                     /* :282 */
    if (!(diff >= -f->trunc)) return 0;                      /* :284 */
    double q = diff / f->trunc;                              /* :286 */
    *dist = q < 1.0 ? q : 1.0;
    *pix = p;
    return 1;
}

static void fill_frame(frame_t *f, const float *cfold, const double *depth, int H, int W,
                       const double intr[4], const double Tinv[16], double trunc)
{
    f->fx = intr[0]; f->fy = intr[1]; f->cx = intr[2]; f->cy = intr[3];
    memcpy(f->T, Tinv, sizeof(f->T));
    f->trunc = trunc; f->H = H; f->W = W; f->depth = depth; f->cfold = cfold;
}

/*
 * Dense grid (TSDFVolume.integrate CPU branch, grid_fusion.py:260-314).
 * Planes are local arrays of shape [nx][Y][Z] (C order); xs[s] is the global x index of
 * local sheet s, so the same routine serves the whole grid, x-slabs and cyclic x-sheets.
 * Returns the number of voxels updated by this frame.
 */
int64_t oracle_dense_integrate(float *tsdf, float *wgt, float *col,
                               const int64_t *xs, int64_t nx, int64_t Y, int64_t Z,
                               const float origin[3], double vs, double trunc,
                               const float *cfold, const double *depth, int H, int W,
                               const double intr[4], const double Tinv[16],
                               double obs, int nthreads)
{
    frame_t f;
    fill_frame(&f, cfold, depth, H, W, intr, Tinv, trunc);
    const float of = (float)obs; /* NEP-50 weak scalar: obs_weight*new_b is an f32 product */
    int64_t updated = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : updated)
    for (int64_t s = 0; s < nx; ++s) {
        double wx = lattice_coord(origin[0], vs, xs[s]);
        for (int64_t j = 0; j < Y; ++j) {
            double wy = lattice_coord(origin[1], vs, j);
            double cxr[3];
            cxr[0] = fma(f.T[1], wy, f.T[0] * wx);
            cxr[1] = fma(f.T[5], wy, f.T[4] * wx);
            cxr[2] = fma(f.T[9], wy, f.T[8] * wx);
            int64_t base = (s * Y + j) * Z;
            for (int64_t k = 0; k < Z; ++k) {
                double wz = lattice_coord(origin[2], vs, k);
                double dist;
                int64_t p;
                if (!project_voxel(&f, cxr, wz, &dist, &p)) continue;
                int64_t i = base + k;
                /* grid_fusion.py:207-211 */
                float w_old = wgt[i];
                float w_new = (float)((double)w_old + obs);
                float prod = w_old * tsdf[i];
                tsdf[i] = (float)(((double)prod + obs * dist) / (double)w_new);
                wgt[i] = w_new;
                /* grid_fusion.py:302-314, f32 throughout */
                float oc = col[i];
                float ob = floorf(oc / 65536.0f);
                float og = floorf((oc - ob * 65536.0f) / 256.0f);
                float orr = oc - ob * 65536.0f - og * 256.0f;
                float nc = f.cfold[p];
                float nb = floorf(nc / 65536.0f);
                float ng = floorf((nc - nb * 65536.0f) / 256.0f);
                float nr = nc - nb * 65536.0f - ng * 256.0f;
                float b = fminf(255.0f, rintf((w_old * ob + of * nb) / w_new));
                float g = fminf(255.0f, rintf((w_old * og + of * ng) / w_new));
                float r = fminf(255.0f, rintf((w_old * orr + of * nr) / w_new));
                col[i] = b * 65536.0f + g * 256.0f + r;
                ++updated;
            }
        }
    }
    return updated;
}

/*
 * Reference hash map viewed as a set of per-voxel entries (hash_fusion.py:103-145 +
 * voxel.py:19-49).  State per voxel is float64 sdf / weight / colour plus a "has entry"
 * flag; the container layout (buckets of 5, offset chains) is irrelevant to the values.
 * obs_weight is NOT a parameter: the reference never forwards it to Voxel.integrate.
 * Returns the number of voxel updates (== get_hash_entry calls) of this frame; *created
 * receives the number of new entries (== add_hash_entry calls).
 */
int64_t oracle_hashvox_integrate(double *sdf, double *wgt, double *col, uint8_t *has,
                                 const int64_t *xs, int64_t nx, int64_t Y, int64_t Z,
                                 const float origin[3], double vs, double trunc,
                                 const float *cfold, const double *depth, int H, int W,
                                 const double intr[4], const double Tinv[16],
                                 int nthreads, int64_t *created)
{
    frame_t f;
    fill_frame(&f, cfold, depth, H, W, intr, Tinv, trunc);
    int64_t updated = 0, made = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : updated, made)
    for (int64_t s = 0; s < nx; ++s) {
        double wx = lattice_coord(origin[0], vs, xs[s]);
        for (int64_t j = 0; j < Y; ++j) {
            double wy = lattice_coord(origin[1], vs, j);
            double cxr[3];
            cxr[0] = fma(f.T[1], wy, f.T[0] * wx);
            cxr[1] = fma(f.T[5], wy, f.T[4] * wx);
            cxr[2] = fma(f.T[9], wy, f.T[8] * wx);
            int64_t base = (s * Y + j) * Z;
            for (int64_t k = 0; k < Z; ++k) {
                double wz = lattice_coord(origin[2], vs, k);
                double dist;
                int64_t p;
                if (!project_voxel(&f, cxr, wz, &dist, &p)) continue;
                int64_t i = base + k;
                if (!has[i]) { /* Voxel(): sdf=1, color=0, weight=0 (voxel.py:8) */
                    has[i] = 1; sdf[i] = 1.0; wgt[i] = 0.0; col[i] = 0.0; ++made;
                }
                /* voxel.py:32-37, obs_weight == 1.0 */
                double w_old = wgt[i], d_old = sdf[i];
                double w_new = w_old + 1.0;
                sdf[i] = (d_old * w_old + dist * 1.0) / w_new;
                wgt[i] = w_new;
                /* voxel.py:39-49: old_* float64, new_* float32 (exact), blend in float64 */
                double oc = col[i];
                double ob = floor(oc / 65536.0);
                double og = floor((oc - ob * 65536.0) / 256.0);
                double orr = oc - ob * 65536.0 - og * 256.0;
                float nc = f.cfold[p];
                float nb = floorf(nc / 65536.0f);
                float ng = floorf((nc - nb * 65536.0f) / 256.0f);
                float nr = nc - nb * 65536.0f - ng * 256.0f;
                double b = fmin(255.0, rint((w_old * ob + (double)nb) / w_new));
                double g = fmin(255.0, rint((w_old * og + (double)ng) / w_new));
                double r = fmin(255.0, rint((w_old * orr + (double)nr) / w_new));
                col[i] = b * 65536.0 + g * 256.0 + r;
                ++updated;
            }
        }
    }
    if (created) *created = made;
    return updated;
}

/* hash_fusion.py:182-190 -- ((x*P1) ^ (y*P2) ^ (z*P3)) floor-mod n, 64-bit, no 32-bit wrap. */
int64_t oracle_hash_function(int64_t x, int64_t y, int64_t z, int64_t table_size)
{
    int64_t h = (x * 73856093LL) ^ (y * 19349669LL) ^ (z * 83492791LL);
    int64_t m = h % table_size;
    if (m != 0 && ((m < 0) != (table_size < 0))) m += table_size; /* np.remainder */
    return m;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
