/*
 * ctg_b200.h -- C-ABI of the B200-native sliced contraction-tree executor.
 *
 * This is the drop-in boundary for cotengra's execution path.  cotengra is pure
 * Python and has no FFI of its own; these entry points are what a binding for
 * that path would call (see INTEGRATION.md for the ctypes stub and the three
 * lines that install it through cotengra's `implementation=` hook).
 *
 * Reference interfaces replaced (paths relative to jcmgray/cotengra @ 2182a79):
 *
 *   ctgb_contract_pair ....... cotengra/contract.py:414 `einsum(eq, a, b)` and
 *                              :521 `tensordot(a, b, axes)` -- one pairwise node,
 *                              lowered there to transpose+reshape -> matmul
 *                              (:364-411); here ONE kernel launch with the index
 *                              permutations folded into the tile loads/stores.
 *   ctgb_reduce_single ....... cotengra/contract.py:332 `_einsum_single` (diag /
 *                              sum / transpose of one operand; preprocessing
 *                              steps :792-796 and single-input trees :797-803).
 *   ctgb_plan_create/execute . cotengra/contract.py:654-837 `Contractor.__call__`
 *                              (the node loop, strip_exponent :816-829) together
 *                              with cotengra/core.py:3943-4030
 *                              `ContractionTree.contract` (slice loop),
 *                              :3775-3819 `slice_key`/`slice_arrays` and
 *                              :3825-3882 `gather_slices` (sum / stack);
 *                              slice_begin/slice_step reproduce the round-robin
 *                              of `contract_mpi` (core.py:4070).
 *
 * Conventions
 *   - All tensors are dense arrays addressed as base + sum(digit * stride), with
 *     strides in ELEMENTS; the caller (host side, cotengra_b200/lowering.py)
 *     turns index labels into strides, so transposes, diagonals (summed strides),
 *     broadcasts (stride 0) and slicing (base offsets) need no data movement.
 *   - Device buffers and the CUDA stream are owned by the caller; the library
 *     owns only immutable plans (plus a small internal descriptor arena).
 *   - Every function returns 0 on success or a CTGB_E_* code; the message is
 *     available from ctgb_last_error() (thread local).
 *   - No CPU fallback exists: without a CUDA device every compute entry point
 *     fails with CTGB_E_CUDA.
 */
#ifndef CTG_B200_H
#define CTG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTGB_ABI_VERSION 1

/* element types (output dtype == input dtype, no casting on the path:
 * cotengra/contract.py has none either) */
enum {
  CTGB_F32 = 0,
  CTGB_F64 = 1,
  CTGB_C64 = 2,
  CTGB_C128 = 3
};

/* status codes; mapped by the Python host to the exceptions the reference raises */
enum {
  CTGB_OK = 0,
  CTGB_E_VALUE = 1,    /* -> ValueError   (bad shapes / descriptor)          */
  CTGB_E_NOTIMPL = 2,  /* -> NotImplementedError                             */
  CTGB_E_CUDA = 3,     /* -> RuntimeError (CUDA failure, or no device)       */
  CTGB_E_MEMORY = 4    /* -> MemoryError  (workspace too small)              */
};

/* Number of int64 words in one pairwise-contraction descriptor.  The word
 * layout is defined in cotengra_b200/csrc/gett_desc.h and mirrored by
 * cotengra_b200/lowering.py (checked at import through ctgb_desc_words()). */
int ctgb_abi_version(void);
int ctgb_desc_words(void);
int ctgb_single_desc_words(void);
const char* ctgb_last_error(void);

/* Device properties the host-side planner sizes grids with. */
int ctgb_device_info(int* sm_count, int* cc_major, int* cc_minor,
                     size_t* smem_optin_bytes);

/* One pairwise node: C = sum_k A * B with arbitrary index placement.
 * `desc` is a host array of ctgb_desc_words() int64 words.  A, B, C are device
 * pointers; `stream` is a cudaStream_t (0 = default stream). */
int ctgb_contract_pair(const int64_t* desc, const void* A, const void* B,
                       void* C, void* stream);

/* One single-operand node: out = diag/sum/transpose of X. */
int ctgb_reduce_single(const int64_t* desc, const void* X, void* out,
                       void* stream);

/* ---- whole-tree plans ---------------------------------------------------- */

typedef struct ctgb_plan ctgb_plan;

/* A tensor slot of the plan.  kind: 0 = network input `input_index` (device
 * pointer supplied at execute time; for sliced inputs the per-slice element
 * offset is sum(digit[slice_pos[j]] * slice_stride[j])), 1 = per-slice
 * workspace at byte offset `offset`, 2 = persistent (slice-invariant)
 * workspace at byte offset `offset`, 3 = the output accumulator. */
typedef struct {
  int32_t kind;
  int32_t input_index;
  int64_t offset;
  int64_t nbytes;
  int32_t n_sliced;         /* sliced indices carried by this input      */
  const int32_t* slice_pos; /* position in the plan's slice-digit list   */
  const int64_t* slice_stride;
} ctgb_tensor;

/* A node of the linear program (cotengra/contract.py:573-651 IR, lowered). */
typedef struct {
  int32_t kind;       /* 0 = pairwise (desc = pair words), 1 = single-operand */
  int32_t a, b, c;    /* tensor slots (b unused for kind 1)                   */
  int32_t invariant;  /* 1: no sliced input below it -> run once per execute  */
  int32_t is_root;    /* 1: writes the output (accumulated over slices)       */
  const int64_t* desc;
} ctgb_node;

typedef struct {
  int32_t dtype;
  int32_t n_inputs;
  int32_t n_tensors;
  const ctgb_tensor* tensors;
  int32_t n_nodes;
  const ctgb_node* nodes;
  /* slicing: mixed-radix digits, most significant first, exactly
   * cotengra/core.py:114-122 + 3775-3800; radix 1 + project >= 0 encodes a
   * projected index (it consumes no digit of the slice id). */
  int32_t n_sliced;
  const int64_t* slice_radix;
  const int64_t* slice_project; /* -1 = sliced normally */
  /* the root's output view: element offset into `out` contributed by the
   * digits of sliced indices that are also OUTPUT indices (gather_slices'
   * stack, core.py:3865-3876); 0 stride for inner sliced indices. */
  const int64_t* slice_out_stride;
  int64_t out_elements;       /* elements of the full output tensor          */
  int64_t workspace_bytes;    /* per-slice arena                              */
  int64_t persistent_bytes;   /* slice-invariant arena                        */
  int32_t strip_exponent;     /* contract.py:816-829 semantics                */
} ctgb_plan_desc;

int ctgb_plan_create(const ctgb_plan_desc* desc, ctgb_plan** plan);
/* strip_exponent only: the single-operand descriptor (ctgb_single_desc_words()
 * words) that maps the dense root result of one slice onto its chunk of the
 * output tensor (identity layout when no sliced index is an output index). */
int ctgb_plan_set_chunk_desc(ctgb_plan* plan, const int64_t* desc);
void ctgb_plan_destroy(ctgb_plan* plan);
size_t ctgb_plan_workspace_bytes(const ctgb_plan* plan);
int64_t ctgb_plan_launches_per_slice(const ctgb_plan* plan);

/* Contract slices slice_begin, slice_begin + slice_step, ... (slice_count of
 * them) and ACCUMULATE their contributions into `out` (device, out_elements of
 * the plan dtype; the caller zeroes it before the first call).  `inputs` is a
 * host array of n_inputs DEVICE pointers to the unsliced, C-contiguous input
 * arrays.  `workspace` must hold ctgb_plan_workspace_bytes() bytes.  With
 * strip_exponent the mantissa is accumulated against the running base-10
 * exponent stored in exponent_dev[0] (device double; core.py:163-170).
 * Asynchronous on `stream`. */
int ctgb_plan_execute(ctgb_plan* plan, const void* const* inputs, void* out,
                      double* exponent_dev, void* workspace,
                      size_t workspace_bytes, int64_t slice_begin,
                      int64_t slice_step, int64_t slice_count, void* stream);

/* Same job with HOST buffers: copies the inputs host->device, runs the slices,
 * copies the accumulated output (and exponent) back, synchronises.  This is the
 * end-to-end call bench.py times as `e2e`.  `workspace` stays a device buffer
 * (it is scratch); input staging memory is taken from its tail. */
int ctgb_plan_execute_host(ctgb_plan* plan, const void* const* host_inputs,
                           const int64_t* input_nbytes, void* host_out,
                           double* host_exponent, void* workspace,
                           size_t workspace_bytes, int64_t slice_begin,
                           int64_t slice_step, int64_t slice_count,
                           void* stream);

/* Per-node device timing for roofline reporting: when enabled, CUDA events are
 * recorded on the execute stream around every node; ctgb_plan_profile_read
 * synchronises and returns the milliseconds of each of the plan's n_nodes
 * nodes for the LAST slice executed (-1 for nodes that did not run). */
int ctgb_plan_profile(ctgb_plan* plan, int enable);
int ctgb_plan_profile_read(ctgb_plan* plan, float* ms, int n_nodes);

/* Measured fp64 tensor-core (DMMA m8n8k4) and fp64 FMA peaks of the current
 * device in TFLOP/s, from a register-resident microbenchmark kernel: the
 * denominators bench.py uses for the fp64 roofline (MEASURED_PEAKS.json holds
 * only HBM and bf16 numbers). */
int ctgb_probe_fp64_peaks(double* dmma_tflops, double* dfma_tflops, void* stream);

/* Number of kernels this library has launched since load (bench.py's
 * `gpu_launches`). */
int64_t ctgb_launch_count(void);
/* ... of which tcgen05 launches whose A tiles are fetched by tensor-map TMA (cp.async.bulk.tensor). */
int64_t ctgb_tensor_map_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* CTG_B200_H */
