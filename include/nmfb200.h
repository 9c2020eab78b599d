/* nmfb200.h - C ABI of libnmfb200.so, the B200-native engine behind the
 * multiplicative-update hot path of colinvaz/nmf-toolbox.
 *
 * The reference is pure MATLAB and has no FFI of its own; its boundary for this
 * path is the MATLAB function-call level.  Every entry point below names the
 * reference interface it replaces (file:line relative to the toolbox root):
 *
 *   nmfb_nmf          [W,H,cost] = nmf(V, num_basis_elems, config)              nmf.m:1
 *   nmfb_cnmf         [W,H,cost] = cnmf(V, num_basis_elems, context_len, config) cnmf.m:1
 *   nmfb_nmfsc        [W,H,cost] = nmfsc(V, num_basis_elems, config)            nmfsc.m:1
 *   nmfb_cnmfsc       [W,H,cost] = cnmfsc(V, num_basis_elems, context_len, config) cnmfsc.m:1
 *   nmfb_lnmf         [W,H,cost] = lnmf(V, num_basis_elems, config)             lnmf.m:1
 *   nmfb_reconstruct  V_hat = ReconstructFromDecomposition(W, H)     ReconstructFromDecomposition.m:1
 *   nmfb_projfunc     [v,usediters] = projfunc(s, k1, k2, nn)                   projfunc.m:1
 *
 * Conventions
 *   - All host matrices are column-major float32, exactly as MATLAB `single`
 *     arrays are laid out: V is m x n, W is m x K (cnmf: m x K x T), H is K x n.
 *   - Host buffers belong to the caller and are never retained; device buffers
 *     for V, W, H and all scratch belong to the handle.
 *   - Calls are synchronous at the boundary: outputs are valid on return.
 *   - A handle is bound to one CUDA device and is not thread-safe (MATLAB calls
 *     from a single interpreter thread).
 *   - Every function returns NMFB_OK (0) or a non-zero nmfb_status; the message
 *     of the last failure is available from nmfb_last_error().  MATLAB error()
 *     sites of the reference map to NMFB_ERR_* codes (listed per function).
 *   - There is no CPU fallback: without a CUDA device nmfb_create fails.
 *   - A call expects the device to itself while it runs.  Several kernels of an
 *     iteration are planned to be resident together and wait for each other on
 *     the device (a Gram product running beside the contraction that consumes it,
 *     helper CTA pairs handing partial sums to the pairs that finish a tile, the
 *     ranks of a multi-GPU run meeting in the sharded W step); the plans size
 *     their grids from the SM count of the device.  Work of another handle or
 *     process that occupies SMs for long can stall such a wait; a wait that lasts
 *     ~2 s traps and the call returns NMFB_ERR_CUDA instead of hanging
 *     (NMFB_TAIL_HELPERS=0 and NMFB_OVERLAP=0 plan without these waits).
 *
 * Multi-GPU (one process per GPU): each rank creates a handle, joins a
 * communicator (nmfb_comm_*), uploads its COLUMN SHARD of V and of H_init and
 * passes the full W_init; W is returned replicated, H as the rank's shard, the
 * cost trace is global.  nmfb_nmf (all divergences, W_fixed and per-source fixed
 * bases included): per iteration every rank fetches its ROW block of the ranks'
 * m x K numerator partials over NVLink peer memory, takes the W step on those
 * rows and delivers them to all ranks; a K x K Gram matrix and a few scalars are
 * all-reduced (without peer access: one NCCL all-reduce of everything).
 * nmfb_cnmf ('euclidean' / 'frobenius'): shards of CONSECUTIVE columns, each at
 * least context_len - 1 wide; the halo columns are exchanged by the engine.
 * nmfb_nmfsc, nmfb_cnmfsc, nmfb_lnmf, nmfb_constrainednmf: one GPU.
 */
#ifndef NMFB200_H_
#define NMFB200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nmfb_handle nmfb_handle;

typedef enum nmfb_status {
  NMFB_OK = 0,
  NMFB_ERR_INVALID_ARGUMENT = 1,   /* null pointers, non-positive sizes, ...                    */
  NMFB_ERR_CUDA = 2,               /* a CUDA / NCCL call failed; see nmfb_last_error            */
  NMFB_ERR_UNSUPPORTED = 3,        /* feature of the reference outside the accelerated path     */
  NMFB_ERR_DIVERGENCE = 4,         /* nmf.m:165-166,196-197 "No update equations defined ..."   */
  NMFB_ERR_AB_ZERO = 5,            /* nmf.m:120-122, cnmf.m:133-135 "alpha = 0 and beta = 0 ..." */
  NMFB_ERR_NEGATIVE_DATA = 6,      /* nmfsc.m:57-59 "Negative values in data!"                  */
  NMFB_ERR_NO_DATA = 7,            /* V has not been set on this handle                         */
  NMFB_ERR_PROJFUNC = 8            /* projfunc produced non-finite values (MATLAB would spin)   */
} nmfb_status;

typedef enum nmfb_divergence {
  NMFB_DIV_EUCLIDEAN = 0,  /* 'euclidean'                      nmf.m:148, cnmf.m:138 */
  NMFB_DIV_KL = 1,         /* 'kl_divergence' | 'kl'           nmf.m:151, cnmf.m:141 */
  NMFB_DIV_FROBENIUS = 2,  /* 'frobenius': cnmf.m only (cnmf.m:138); same updates as
                              euclidean but the reference's cost switch has no such
                              case, so cost = sparsity terms only (cnmf.m:239-251).
                              nmf.m rejects it (NMFB_ERR_DIVERGENCE).                */
  NMFB_DIV_IS = 3,         /* 'is_divergence' | 'is'  nmf.m:154-156,185-187,211-212 (nmfb_nmf, also
                              column-sharded; nmfb_cnmf, cnmf.m:144-146)              */
  NMFB_DIV_AB = 4          /* 'ab_divergence' | 'ab'  nmf.m:157-164,188-195,213-214 with config
                              alpha, beta; alpha == 0 selects the dual updates
                              (nmf.m:124-128); same availability as NMFB_DIV_IS       */
} nmfb_divergence;

/* Accuracy of the cost trace and of the stop test (nmf.m:221-224: stop when 0 < cost(i-1) - cost(i) < tolerance).
 * The contractions multiply tf32-rounded operands with fp32 accumulation, so every cost entry agrees with the
 * float64 reference to about 1e-5 relative (measured <= 1.3e-5; the trace form subtracts terms ~8x the cost, the
 * direct form is ~3e-6).  A `tolerance` below ~1e-5 * cost therefore cannot be resolved: the loop may stop one or
 * more iterations away from the reference's iteration, or run to maxiter.  For resolvable tolerances the default mode
 * stops within one iteration of the reference, NMFB_COST_DIRECT at the same iteration (tests/test_gpu_parity.py). */
typedef enum nmfb_cost_mode {
  NMFB_COST_AUTO = 0,   /* Euclidean: Gram/trace identity (no extra pass over V); KL: fused */
  NMFB_COST_DIRECT = 1  /* Euclidean: explicit 0.5*sum((V - W*H).^2) each iteration        */
} nmfb_cost_mode;

/* Mirror of the reference's `config` struct (nmf.m:17-65, cnmf.m:28-75,
 * nmfsc.m:11-32).  Zero-initialise, then set what you need:
 *   W_init/H_init NULL  -> uniform random (seeded by `seed`), as rand() in
 *                          nmf.m:277,298 / cnmf.m:331 / nmfsc.m:74,79
 *   maxiter   <= 0      -> 100    (nmf.m:404-406)
 *   tolerance <= 0      -> 1e-3   (nmf.m:409-411)
 *   *_sparsity < 0      -> 0      (nmf.m:321-333); nmfsc clamps to <= 1 (nmfsc.m:90,103) */
typedef struct nmfb_config {
  int divergence;        /* nmfb_divergence (ignored by nmfb_nmfsc)                     */
  double alpha, beta;    /* NMFB_DIV_AB only (nmf.m:23-28); both 0 -> NMFB_ERR_AB_ZERO  */
  const float* W_init;   /* m x K (x T) column-major, or NULL                           */
  const float* H_init;   /* K x n_local column-major, or NULL                           */
  double W_sparsity;     /* lambda_W (nmf/cnmf) or Hoyer sparseness of W columns (nmfsc) */
  double H_sparsity;     /* lambda_H (nmf/cnmf) or Hoyer sparseness of H rows (nmfsc)    */
  int W_fixed, H_fixed;  /* nmf.m:51-60                                                 */
  int maxiter;
  double tolerance;
  unsigned long long seed; /* only used when an init pointer is NULL                    */
  int cost_mode;         /* nmfb_cost_mode                                              */
  /* Optional per-basis overrides for nmfb_nmf (NULL = the scalars above apply to every basis).
   * They carry the reference's multi-source convention (cell arrays, nmf.m:11-16,51-60,
   * 284-400) across the C boundary: the caller concatenates the sources' bases and passes, for
   * every basis column k < K, the setting of the source it belongs to.  Exact, because the
   * per-source loops of nmf.m:144-171 / 175-201 never refresh V_hat between sources.       */
  const double* W_sparsity_k; /* K values of lambda_W (negative -> 0)                      */
  const double* H_sparsity_k; /* K values of lambda_H                                      */
  const int* W_fixed_k;       /* K flags: basis k of W is held fixed (nmf.m:145)           */
  const int* H_fixed_k;       /* K flags: row k of H is held fixed (nmf.m:176)             */
} nmfb_config;

/* ---- handle ------------------------------------------------------------- */
int nmfb_create(nmfb_handle** out, int device);
void nmfb_destroy(nmfb_handle* h);
const char* nmfb_last_error(const nmfb_handle* h); /* h may be NULL: last create error */

/* The handle keeps the device blocks of finished calls for reuse by the next call of the same
 * shape (cudaMalloc / cudaFree of GiB-sized blocks cost as much as ~100 iterations); at most a
 * third of the device memory, released on allocation failure, by nmfb_destroy and by this call.
 * NMFB_NO_POOL=1 in the environment disables the cache. */
int nmfb_trim(nmfb_handle* h);

/* Upload V (m x n column-major float32, host memory; pinned or pageable). */
int nmfb_set_V(nmfb_handle* h, const float* V_host, int m, int n);
/* Adopt a V that already lives on the handle's device (column-major, leading
 * dimension ld >= m, ld % 4 == 0, 16-byte aligned).  Not copied, not modified;
 * must stay alive until the next nmfb_set_V* or nmfb_destroy. */
int nmfb_set_V_device(nmfb_handle* h, const float* V_dev, int m, int n, long long ld);

/* ---- the reference's functions ------------------------------------------ */
/* W_out: m*K floats, H_out: K*n floats, cost_out: >= maxiter doubles (after
 * defaulting).  *n_cost = number of executed iterations = length of the
 * reference's trimmed cost vector (nmf.m:222).  Outputs may be NULL. */
int nmfb_nmf(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out,
             double* cost_out, int* n_cost);
/* [W, H, cost] = lnmf(V, num_basis_elems, config)  (lnmf.m:1).
 * KL-type updates with unit-sum bases and H <- sqrt(H .* W'(V./V_hat)) (lnmf.m:63-83) on the fused
 * KL kernels (num_basis_elems <= 128; the unfused contractions beyond that; one GPU).  config: W_init, H_init, W_fixed, H_fixed, maxiter,
 * tolerance (divergence and sparsity fields are ignored: lnmf.m has none).  As in the reference the
 * cost vector is NOT trimmed when the loop stops early (lnmf.m:88-90): *n_cost == maxiter, entries
 * after the stopping iteration are 0. */
int nmfb_lnmf(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out, double* cost_out,
              int* n_cost);

/* [W, H, Z, A, cost] = constrainednmf(V, labels, num_basis_elems, config)  (constrainednmf.m:1; SURVEY 8f
 * item 4).  Label-constrained NMF V ~ W*Z*A: the W step is nmf.m's (constrainednmf.m:185-209), the encoding is
 * H = Z*A with A the 0/1 label-indicator matrix, and Z takes the multiplicative step on the gradients summed
 * over the samples of a class (213-237).  The label handling of constrainednmf.m:147-170 is host work and
 * stays with the caller (nmf_toolbox_b200/api.py, matlab/constrainednmf.m): V must be set with its columns in
 * the reference's ORDERED arrangement (unlabeled samples first, then the classes, each contiguous), and
 * col2z[j] in [0, nz) (non-decreasing, every value used) names the column of Z that sample j reads, i.e. the
 * row of A holding its 1.  config: divergence / alpha / beta, W_init, W_sparsity, W_fixed, maxiter, tolerance
 * as for nmfb_nmf; H_sparsity and H_fixed carry the reference's Z_sparsity and Z_fixed; H_init is ignored.
 * Z_init (K x nz column-major, may be NULL = uniform random) is an EXTENSION: the reference draws Z = rand(...)
 * unconditionally (line 174), which no test could reproduce.  H_out (K x n) is in the ordered arrangement.
 * The non-dual alpha-beta branch of the reference's Z update multiplies mismatched matrices (line 229) and
 * fails unless m == num_basis_elems: NMFB_ERR_UNSUPPORTED.  One GPU. */
int nmfb_constrainednmf(nmfb_handle* h, int K, const nmfb_config* cfg, const int* col2z, int nz,
                        const float* Z_init, float* W_out, float* H_out, float* Z_out, double* cost_out,
                        int* n_cost);

/* W_out: m*K*T floats (m x K x T column-major). */
int nmfb_cnmf(nmfb_handle* h, int K, int T, const nmfb_config* cfg, float* W_out, float* H_out,
              double* cost_out, int* n_cost);
/* cost_out: >= maxiter+1 doubles; cost_out[0] is the initial cost (nmfsc.m:137-139).
 * W_out, H_out factor V/max(V) (nmfsc.m:62). */
int nmfb_nmfsc(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out,
               double* cost_out, int* n_cost);

/* [W, H, cost] = cnmfsc(V, num_basis_elems, context_len, config)  (cnmfsc.m:1).
 * Convolutive NMF with Hoyer sparseness on H: projected-gradient H step with the reference's line
 * search when config.H_sparsity > 0 (cnmfsc.m:166-199), multiplicative H step with row
 * normalisation otherwise (202-209), frame-by-frame multiplicative W step (257-263).  V is rescaled
 * by its maximum (line 72), so W, H factor V / max(V).  cost_out needs maxiter + 1 entries
 * (cost(1) = initial objective).  config.W_sparsity > 0 follows the reference literally, quirks
 * included (cnmfsc.m:93-111, 218, 235): the W line search reconstructs its trial from one frame
 * alone and on ordinary data ends by step-size underflow with the cost trimmed (lines 245-249).
 * One GPU. */
int nmfb_cnmfsc(nmfb_handle* h, int K, int T, const nmfb_config* cfg, float* W_out, float* H_out, double* cost_out,
                int* n_cost);
/* V_hat (m x n, host) = W*H, or sum_t W(:,:,t) * shift(H, t-1) when T > 1.
 * Independent of any V set on the handle. */
int nmfb_reconstruct(nmfb_handle* h, const float* W, const float* H, int m, int K, int T, int n,
                     float* Vhat_out);
/* Hoyer projection of `count` vectors of length N stored back to back
 * (projfunc.m:13-55); iters_out (may be NULL) receives usediters per vector. */
int nmfb_projfunc(nmfb_handle* h, const float* s, int N, int count, double k1, double k2, int nn,
                  float* v_out, int* iters_out);

/* ---- stepping interface (what nmfb_nmf does internally; used by bench.py to
 *      time exactly K iterations with V, W, H resident) ---------------------- */
int nmfb_nmf_begin(nmfb_handle* h, int K, const nmfb_config* cfg);
/* Enqueue `iters` more iterations on the handle's stream (asynchronous). */
int nmfb_nmf_step(nmfb_handle* h, int iters);
/* Wait for the queued iterations; returns how many have been executed so far
 * and the device time (CUDA events on the handle's stream) they took in total. */
int nmfb_nmf_sync(nmfb_handle* h, int* iters_done, double* device_ms);
int nmfb_nmf_end(nmfb_handle* h, float* W_out, float* H_out, double* cost_out, int* n_cost);
/* Optional per-kernel timing (CUDA events on the handle's stream around every
 * launch of the two large contractions of the Euclidean nmf iteration). */
int nmfb_profile_enable(nmfb_handle* h, int on);
int nmfb_profile_get(nmfb_handle* h, double* ms_w_gemm, double* ms_h_gemm, int* count);
/* ms_out[5], averages per launch group.  nmf euclidean: W-step GEMM, H-step GEMM, gram(H)+cost,
 * element-wise W step, gram(W).  nmf KL: [0] fused W half, [1] fused H half.  cnmf: [0] A/B GEMM,
 * [1] P/D GEMM (+ slab sum), [2] fold + H update.  nmfsc: [0] gradient GEMMs of the H step, [1] one
 * objective evaluation, [2] step + projfunc + split of one trial, [3] W step. */
int nmfb_profile_get_all(nmfb_handle* h, double* ms_out);
/* Iterations executed by, and device time (CUDA events on the handle's stream, ms) of, the iteration
 * loop of the last nmfb_nmf / nmfb_lnmf / nmfb_cnmf / nmfb_nmfsc / nmfb_cnmfsc call on this handle:
 * setup, uploads and downloads excluded (bench.py's resident timing of the one-call algorithms). */
int nmfb_last_loop(nmfb_handle* h, int* iters, double* device_ms);
/* nmfb_nmfsc: how often the line searches of the last call halved their step (nmfsc.m:169, 220):
 * out[2*i] for the H search, out[2*i+1] for the W search of iteration i.  Returns the number of
 * entries available (2 x executed iterations); at most `capacity` are written. */
int nmfb_last_halvings(nmfb_handle* h, int* out, int capacity);
/* Number of kernel launches issued by the handle since creation. */
long long nmfb_launch_count(const nmfb_handle* h);
/* Number of cudaMalloc calls the handle has made (allocations that missed its block cache): constant
 * across repeated calls of the same shape, i.e. a steady-state call allocates nothing. */
long long nmfb_malloc_count(const nmfb_handle* h);

/* ---- multi-GPU ---------------------------------------------------------- */
#define NMFB_UNIQUE_ID_BYTES 128
int nmfb_comm_unique_id(char id_out[NMFB_UNIQUE_ID_BYTES]);
int nmfb_comm_init(nmfb_handle* h, const char id[NMFB_UNIQUE_ID_BYTES], int rank, int nranks);

const char* nmfb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NMFB200_H_ */
