// launch.h -- host-side launch wrappers shared by the .cu translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct FrameParams;
struct VolumeParams;
struct CullParams;
struct BrickRef;

// Reset by one cudaMemsetAsync at the start of every frame.
struct FrameScratch {
    unsigned int max_depth_bits; // float bits of the largest valid depth (rounded up)
    int vis_count;               // entries in the visible-brick list
    unsigned long long updated;  // voxels updated by this frame
    int cand_count;              // hash: candidate keys emitted by the ray walk (after dedup)
    int new_keys;                // hash: keys inserted this frame
    int error;                   // sticky per-frame error flag (tsdf_status), 0 = ok
    int pad;
};

// Lives for the lifetime of the handle.
struct Totals {
    unsigned long long updated;      // voxels updated since create
    unsigned long long last_visited; // bricks/blocks visited by the last frame
    unsigned long long max_probe;    // hash: longest probe sequence
    unsigned long long error;        // hash: sticky error (tsdf_status) raised inside a kernel
};

void launch_fill_bricks(float *bricks, size_t nbricks, int sm_count, cudaStream_t s);
void launch_prep_frame(const uint8_t *rgb, const double *depth, uint32_t *cfold, FrameScratch *fs, int npix,
                       int sm_count, cudaStream_t s);
void launch_cull_bricks(const FrameParams &f, const CullParams &c, int nlbx, int nby, int nbz, int rank, int world,
                        BrickRef *list, FrameScratch *fs, cudaStream_t s);
int integrate_dense_ctas_per_sm();
void launch_integrate_dense(const FrameParams &f, const VolumeParams &v, const BrickRef *list, float *bricks,
                            FrameScratch *fs, Totals *tot, int grid, cudaStream_t s);
void launch_unbrick(const float *bricks, int n_x, int dim_y, int dim_z, int nby, int nbz, float *tsdf, float *wgt,
                    float *col, int sm_count, cudaStream_t s);
