/*
 * prs_b200.h - C-ABI of the B200-native kd-tree neighbour resampling path.
 *
 * The reference path (pyresample/kd_tree.py) has no FFI of its own: it is a plain
 * Python module whose native work happens inside pykdtree (k-NN), pyproj/PROJ
 * (inverse projection) and numpy (ECEF transform, masks, gather, weighting).
 * This library replaces exactly those native pieces.  Each entry point names
 * the reference interface it stands in for.  All pointers are DEVICE pointers
 * unless marked "host"; sizes are element counts; every function returns 0 on
 * success or a negative PRS_E* code, and prs_last_error() gives the message.
 * Work is enqueued on `stream` (a cudaStream_t passed as void*); functions
 * that must return a value to the host synchronise that stream themselves.
 *
 * INTEGRATION.md shows the ctypes binding a pyresample maintainer would add.
 */
#ifndef PRS_B200_H
#define PRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRS_ABI_VERSION 1

/* coordinate / tree dtype (SwathDefinition.dtype, pyresample/geometry.py:662-666) */
#define PRS_F32 0
#define PRS_F64 1

/* error codes */
#define PRS_OK 0
#define PRS_EINVAL (-1)  /* bad argument */
#define PRS_ECUDA (-2)   /* CUDA runtime error (message has the cudaError string) */
#define PRS_ENOSUP (-3)  /* unsupported parameter combination */
#define PRS_ENOMEM (-4)

/* projection kinds (subset of PROJ needed by the path) */
#define PRS_PROJ_LONGLAT 0
#define PRS_PROJ_EQC 1
#define PRS_PROJ_MERC 2
#define PRS_PROJ_LAEA 3
#define PRS_PROJ_STERE 4
#define PRS_PROJ_GEOS 5

#define PRS_MODE_N_POLE 0
#define PRS_MODE_S_POLE 1
#define PRS_MODE_OBLIQ 2
#define PRS_MODE_EQUIT 3

/* Inverse-projection parameters, prepared on the host (pyresample_b200/proj.py).
 * Stands in for the pyproj.Transformer built at pyresample/geometry.py:2630-2632. */
typedef struct prs_proj_params {
    int32_t kind;        /* PRS_PROJ_* */
    int32_t mode;        /* PRS_MODE_* for laea / stere */
    int32_t ellipsoidal; /* es != 0 */
    int32_t flip_axis;   /* geos: sweep == x */
    double a, ra, es, e, one_es;
    double lam0, phi0;   /* radians */
    double x0, y0, to_meter;
    double pm_deg;       /* +pm, degrees: lon/lats are reported on the Greenwich geodetic CRS */
    int32_t over;
    int32_t _pad;
    double c[12];        /* per-projection constants, see pyresample_b200/proj.py */
} prs_proj_params;

/* Pixel-centre generator of an AreaDefinition (pyresample/geometry.py:1374-1380,1593-1609):
 *   x_j = arange(width, dtype)[j] *  pixel_size_x + pixel_upper_left_x
 *   y_i = arange(height, dtype)[i] * -pixel_size_y + pixel_upper_left_y      (evaluated in `dtype`) */
typedef struct prs_area_grid {
    int64_t width, height;
    double pixel_size_x, pixel_size_y;
    double pixel_upper_left_x, pixel_upper_left_y;
} prs_area_grid;

/* Result of data_reduce._get_valid_index's boundary analysis (pyresample/data_reduce.py:258-305);
 * the O(boundary) part runs on the host, the O(N) compares on the device. */
#define PRS_REDUCE_ALL 0      /* keep everything */
#define PRS_REDUCE_LAT_GE 1   /* round(angle_sum) == -360: lat >= lat_min */
#define PRS_REDUCE_LAT_LE 2   /* round(angle_sum) == +360: lat <= lat_max */
#define PRS_REDUCE_BOX 3      /* lat band and lon band */
#define PRS_REDUCE_BOX_DATELINE 4 /* lat band and (lon >= lon_min or lon <= lon_max) */
typedef struct prs_reduce_box {
    int32_t kind;
    int32_t _pad;
    double lat_min, lat_max, lon_min, lon_max; /* already buffered; compared in double after promotion */
} prs_reduce_box;

/* A set of query (target) points: explicit lon/lat arrays or rows of an area. */
#define PRS_TARGET_LONLAT 0
#define PRS_TARGET_AREA 1
typedef struct prs_target {
    int32_t kind;
    int32_t dtype;            /* dtype the lon/lats are produced in (= source dtype, kd_tree.py:513-514) */
    const void *lon, *lat;    /* PRS_TARGET_LONLAT: n values each */
    int64_t n;                /* number of target points (area: nrows * width) */
    prs_proj_params proj;     /* PRS_TARGET_AREA */
    prs_area_grid grid;
    int64_t row0, nrows;      /* rows of the area owned by this call: the multi-GPU / `segments` shard axis.
                                 Local row l is global row  row0 + (l / row_block) * row_stride + l % row_block ;
                                 row_block == 0 means one contiguous block [row0, row0 + nrows). */
    const uint8_t *valid_extra; /* optional extra validity AND-ed in (grid -> swath reduce, kd_tree.py:438-451) */
    int64_t row_block, row_stride; /* block-cyclic row ownership (rank r of N, stripes of B rows: row0 = r*B,
                                      row_block = B, row_stride = N*B) */
} prs_target;

/* cells per cube-face side of the target coverage mask (prs_target_coverage / prs_index_build `cover`) */
#define PRS_COVER_CELLS 512

typedef struct prs_index prs_index; /* opaque spatial index over the valid source points */

int prs_abi_version(void);
const char *prs_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (measurement aid for bench.py) */
int64_t prs_launch_count(void);

/* ---- coordinate transform --------------------------------------------------------------------------
 * Cartesian.transform_lonlats (pyresample/_spatial_mp.py:155-169), lonlat2xyz
 * (future/resamplers/_transform_utils.py:22-33), BaseDefinition.get_cartesian_coords (geometry.py:465-516).
 * xyz_out: n*3 values of `dtype`, AoS, R = 6370997.0. */
int prs_lonlat_to_xyz(const void *lon, const void *lat, int64_t n, int dtype, void *xyz_out, void *stream);

/* AreaDefinition.get_lonlats(data_slice=, dtype=) (geometry.py:2558-2645) for the strided pixel rectangle
 * rows row0 + i*row_step (i < nrows), cols col0 + j*col_step (j < ncols); outputs nrows*ncols, row-major.
 * Either output may be NULL.  xyz_out (optional) receives the ECEF transform of the same pixels. */
int prs_area_lonlats(const prs_proj_params *proj, const prs_area_grid *grid, int dtype,
                     int64_t row0, int64_t row_step, int64_t nrows, int64_t col0, int64_t col_step, int64_t ncols,
                     void *lon_out, void *lat_out, void *xyz_out, void *stream);

/* The same for the WHOLE area as ONE RANK of a row-sharded run sees it: 16 x 16 pixel tiles that cannot hold a pixel
 * within the coverage mask `cover` (prs_target_coverage of the rank's rows) are not transformed and come out as NaN,
 * i.e. invalid.  The index built from these arrays with the same `cover` holds exactly the records it would hold
 * from prs_area_lonlats' output; only the validity mask / the count of valid points differ, so this form is for the
 * fused nearest path, which consumes neither. */
int prs_area_lonlats_covered(const prs_proj_params *proj, const prs_area_grid *grid, int dtype, const uint8_t *cover,
                             void *lon_out, void *lat_out, void *stream);

/* ---- validity masks -------------------------------------------------------------------------------
 * _get_valid_input_index / _get_valid_output_index range test (kd_tree.py:406,454) AND the reduce_data box
 * (data_reduce.py:282-305).  `premask` (optional, 1 = masked out) carries np.ma masks (kd_tree.py:426-428).
 * valid_out: n bytes (0/1).  *n_valid_host (host, optional) receives the number of valid points (syncs). */
int prs_valid_mask(const void *lon, const void *lat, int64_t n, int dtype, const prs_reduce_box *box /*host, optional*/,
                   const uint8_t *premask, uint8_t *valid_out, int64_t *n_valid_host, void *stream);

/* ---- neighbour index ------------------------------------------------------------------------------
 * Stands in for pykdtree.kdtree.KDTree(input_coords) (kd_tree.py:464-489; xarray path kd_tree.py:970-981).
 * Builds the cell index over the points with valid[i] != 0.  `exclude` (optional, 1 = skip) is pykdtree's
 * query `mask=` (future/resamplers/nearest.py:56-67): excluded points keep their rank but are never returned.
 * tree_dtype is the dtype the ECEF coordinates and distances are computed in (kd_tree.py:536-542; the xarray
 * path always uses PRS_F64, kd_tree.py:979).  k_hint / radius_hint size the cells. */
int prs_index_build(const void *lon, const void *lat, int64_t n, int dtype, uint8_t *valid,
                    const prs_reduce_box *box /*host, optional*/, const uint8_t *premask,
                    const uint8_t *exclude, int tree_dtype, int k_hint, double radius_hint,
                    const uint8_t *cover, int flags, prs_index **index_out, void *stream);
/* flags for prs_index_build */
#define PRS_BUILD_RANKS 1 /* keep the rank table: needed by prs_query (index_array is in rank space) when some
                             source points are invalid; the fused resample entry points do not need it */
#define PRS_BUILD_COMPUTE_VALID 2 /* `valid` is an OUTPUT (may be NULL): the validity of prs_valid_mask (range test,
                             `box`, `premask`) is evaluated inside the build instead of being read from `valid` */
/* `cover` (optional, from prs_target_coverage): source points whose coverage cell is 0 cannot be within the
 * radius of any target point of this rank and are left out of the index (they keep their rank). */

/* Coverage mask of a target (one rank's rows) at PRS_COVER_CELLS^2 cells per cube face: cell = 1 when some
 * target point may have a source point of that cell within `radius` (conservative, dilated by one cell).
 * cover_out: uint8 [6 * PRS_COVER_CELLS^2], zeroed by the call. */
int prs_target_coverage(const prs_target *target, double radius, uint8_t *cover_out, void *stream);
int prs_index_destroy(prs_index *index);
/* info[0]=n_source, [1]=n_valid (pykdtree .n), [2]=n_indexed, [3]=cells per cube-face side, [4]=bytes held,
 * [5]=tree dtype, [6]=coarse cells per side, [7]=1 if every source point is valid */
int prs_index_info(const prs_index *index, int64_t info[8]);

/* ---- query ----------------------------------------------------------------------------------------
 * Stands in for KDTree.query(coords, k, eps=0, distance_upper_bound=radius) + the target-side work of
 * _query_resample_kdtree (kd_tree.py:492-550): target lon/lat generation, validity, ECEF, k-NN.
 * Outputs cover ALL target points in raveled order (the host compacts by valid_out when some are invalid):
 *   idx_out   uint32 [n * k]  rank among valid source points, n_valid where there is no neighbour
 *   dist_out  tree dtype [n * k]  chord distance in metres, +inf where there is no neighbour (may be NULL)
 *   valid_out uint8 [n]  valid_output_index (may be NULL)
 *   flags_host (host, optional): [0] = 1 if any k-th neighbour was found (kd_tree.py:380-384 warning),
 *                                [1] = number of invalid target points. Reading them synchronises the stream.
 * Ties in distance are broken to the lowest source index; candidates must satisfy d2 < (dtype)(radius^2). */
int prs_query(const prs_index *index, const prs_target *target, int k, double radius,
              uint32_t *idx_out, void *dist_out, uint8_t *valid_out, int64_t *flags_host, void *stream);

/* The library keeps memory between calls: a grow-only workspace for the index build's temporaries (17 bytes per
 * source point of the largest build so far, per device) and whatever the device's stream-ordered pool has cached.
 * prs_release_cached_memory() waits for the current device, frees the workspace and trims the pool; live indexes
 * are not touched.  Long-running services call it after an unusually large job. */
int prs_release_cached_memory(void);

/* ---- fused query + sample gather --------------------------------------------------------------------
 * resample_nearest (kd_tree.py:64-110) = get_neighbour_info + get_sample_from_neighbour_info('nn') without
 * materialising the neighbour arrays.  data: source rows of row_bytes bytes each, indexed by RAVELED SOURCE
 * index (not compacted); out: n target rows; fill_row (device, row_bytes) is written where no neighbour
 * exists or the target point is invalid (kd_tree.py:705-723). */
int prs_resample_nearest(const prs_index *index, const prs_target *target, double radius,
                         const void *data, int64_t row_bytes, const void *fill_row, void *out, void *stream);

/* Staged form of prs_resample_nearest for area targets, for callers that overlap host<->device transfers with
 * the kernels (pyresample_b200/kd_tree.py:_nearest_pipelined): the output of tiles that cannot have a neighbour
 * does not depend on `data`, so it can be produced - and read back - while `data` is still being uploaded.
 *   prs_resample_scratch_bytes: size of the device scratch buffer the two stages share (0: not an area target);
 *   prs_resample_nearest_fill:   classifies the tiles (32 columns x PRS_TILE_ROWS rows), writes fill_row into every
 *                                pixel of the tiles without neighbours, lists the others in scratch and sets
 *                                tile_row_live[r] = 1 (device uint8 [ceil(nrows / PRS_TILE_ROWS)], optional) for each
 *                                band of PRS_TILE_ROWS rows that holds such a tile - the other bands are final;
 *   prs_resample_nearest_search: search + gather (kd_tree.py:545-548, 705-723) for the tiles listed in scratch.
 * Both stages together write exactly what prs_resample_nearest writes. */
#define PRS_TILE_ROWS 8
int64_t prs_resample_scratch_bytes(const prs_target *target);
int prs_resample_nearest_fill(const prs_index *index, const prs_target *target, double radius, int64_t row_bytes,
                              const void *fill_row, void *out, void *scratch, uint8_t *tile_row_live, void *stream);
int prs_resample_nearest_search(const prs_index *index, const prs_target *target, double radius, const void *data,
                                int64_t row_bytes, const void *fill_row, void *out, void *scratch, void *stream);

/* resample_gauss (kd_tree.py:113-189) fused: w = exp(-d^2 / sigma_c^2), _resample_with_weights
 * (kd_tree.py:741-818) and optionally _calculate_uncertainty (kd_tree.py:821-859).
 * data: [n_source * channels] of data_dtype (PRS_F32/PRS_F64), out/stddev/count: [n * channels] of the
 * promoted dtype (f32 only when both tree dtype and data dtype are f32), stddev/count optional.
 * sigmas: host array of `channels` doubles.  fill: value for pixels with no weight (kd_tree.py:807-809). */
int prs_resample_gauss(const prs_index *index, const prs_target *target, int k, double radius,
                       const void *data, int data_dtype, int channels, const double *sigmas_host, double fill,
                       void *out, void *stddev_out, void *count_out, int64_t *flags_host, void *stream);

/* ---- sampling from precomputed neighbour info -------------------------------------------------------
 * get_sample_from_neighbour_info (kd_tree.py:566-652) on arrays the caller already holds.
 * rank_to_source: uint32 [n_valid] raveled source index of each valid rank (NULL = identity).
 * index: uint32 [n_out * k] ranks with sentinel n_valid. */
int prs_gather_nearest(const uint32_t *index, int64_t n_out, uint32_t n_valid, const uint32_t *rank_to_source,
                       const void *data, int64_t row_bytes, const void *fill_row, void *out, void *stream);

/* Weighted sum with explicit weights or Gaussian weights computed from distances.
 * weights: [n_planes][n_out * k] of weight_dtype, channel c uses plane plane_of_channel[c] (host array);
 * pass weights = NULL and sigmas_host != NULL to compute Gaussian weights from dist (dist dtype = weight_dtype). */
int prs_gather_weighted(const uint32_t *index, const void *dist, int64_t n_out, int k, uint32_t n_valid,
                        const uint32_t *rank_to_source, const void *data, int data_dtype, int channels,
                        const void *weights, int weight_dtype, int n_planes, const int32_t *plane_of_channel_host,
                        const double *sigmas_host, double fill,
                        void *out, void *stddev_out, void *count_out, void *stream);

/* What the reference hands to a weight function (kd_tree.py:751-768): the (n_out, k) distances as float64 with 1 in
 * place of a missing neighbour (index >= n_valid).  n = n_out * k entries; dist_dtype PRS_F32 / PRS_F64. */
int prs_weight_distances(const uint32_t *index, const void *dist, int dist_dtype, int64_t n, uint32_t n_valid,
                         double *out, void *stream);

/* ---- helpers ----------------------------------------------------------------------------------------
 * Row compaction by a byte mask (numpy boolean indexing `a[valid]`, kd_tree.py:469-470,530-531):
 * copies the rows with mask != 0, in order; returns the number kept through *n_kept_host (syncs). */
int prs_compact_rows(const uint8_t *mask, int64_t n, int64_t row_bytes, const void *in, void *out,
                     int64_t *n_kept_host, void *stream);
/* rank -> raveled source index table for a validity mask (np.flatnonzero), out: uint32 [n_valid] */
int prs_rank_to_source(const uint8_t *valid, int64_t n, uint32_t *out, int64_t *n_valid_host, void *stream);
/* scatter rows into the positions where mask != 0 (full_result[valid_output_index] = result, kd_tree.py:719-723);
 * positions with mask == 0 receive fill_row. */
int prs_scatter_rows(const uint8_t *mask, int64_t n, int64_t row_bytes, const void *in, const void *fill_row,
                     void *out, void *stream);

/* ---- bilinear resampling on the same index (pyresample/bilinear/_base.py) ------------------------------------
 * BilinearBase.get_bil_info (:99-119) = a k = 32 prs_query + the two calls below; get_sample_from_bil_info (:251-262)
 * = prs_bilinear_sample.
 * prs_project_forward: `pyproj.Proj(target.proj_str)(lon, lat)` (_base.py:188-191, 299-307) for every source point;
 *   xy_out: double [n * 2] (x, y) pairs, (inf, inf) where the projection fails.
 * prs_bilinear_info: per target pixel of the area `grid` (n_t = width * height, raveled) the closest neighbour in
 *   each quadrant among the k neighbours `index` (ranks from prs_query, sentinel n_valid, sorted by distance;
 *   _base.py:391-411, 573-611) and the fractional distances (_base.py:414-570).
 *   t_out, s_out: double [n_t] (NaN where no enclosing quadrilateral exists); idx4_out: uint32 [n_t * 4] ranks of the
 *   upper-left, upper-right, lower-left, lower-right corner (rank of the first neighbour where a quadrant is empty).
 * prs_bilinear_sample: p1 (1-s)(1-t) + p2 s (1-t) + p3 (1-s) t + p4 s t (_base.py:653-660) in float64 for every
 *   channel; NaN results become `fill` unless fill is NaN (_numpy_resampler.py:276-282).  out: double [n_out * channels]. */
int prs_project_forward(const void *lon, const void *lat, int64_t n, int dtype, const prs_proj_params *proj,
                        void *xy_out, void *stream);
int prs_bilinear_info(const uint32_t *index, int64_t n_t, int k, uint32_t n_valid, const uint32_t *rank_to_source,
                      const void *xy, const prs_area_grid *grid, double *t_out, double *s_out, uint32_t *idx4_out,
                      void *stream);
int prs_bilinear_sample(const uint32_t *idx4, const double *t, const double *s, int64_t n_out,
                        const uint32_t *rank_to_source, const void *data, int data_dtype, int channels, double fill,
                        double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PRS_B200_H */
