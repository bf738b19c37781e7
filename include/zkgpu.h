/* zkgpu.h — C ABI of the B200-native STARK proving path for zk_evm's evm_arithmetization.
 *
 * The reference (Rust) has no FFI layer; its seams are generic call sites into plonky2/starky.  Each entry point
 * below names the reference interface it replaces (paths relative to /root/reference):
 *
 *   S1  PolynomialBatch::from_values / from_coeffs      evm_arithmetization/src/prover.rs:100-107, verifier.rs:69-76
 *   S2  starky get_ctl_data                             evm_arithmetization/src/prover.rs:134-144
 *   S3  starky prove_with_commitment (prove_single_table)  evm_arithmetization/src/prover.rs:301-341
 *   top prove_with_traces / prove_with_commitments      evm_arithmetization/src/prover.rs:72-194, 211-293
 *
 * Conventions: every function returns ZKGPU_OK (0) or a negative zkgpu_status; nothing throws or aborts across
 * the boundary; zkgpu_last_error() gives the message of the last failure on the calling thread.  Field elements
 * are canonical Goldilocks u64 (< 2^64 - 2^32 + 1), little-endian in memory.  Columns are column-major: one
 * contiguous array of n u64 per column, exactly `PolynomialValues<F>::values`.  Input pointers are borrowed for
 * the duration of the call and may be pageable or pinned host memory (or device memory when the
 * ZKGPU_MEM_DEVICE flag is given); outputs are written to caller-owned buffers or returned as opaque handles
 * released with the matching *_free.  One zkgpu_ctx per (host thread, device); calls on a ctx are serialised
 * on its stream.  The library fails loudly (ZKGPU_ERR_CUDA) when no CUDA device is usable: there is no CPU path.
 */
#ifndef ZKGPU_H
#define ZKGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ZKGPU_OK = 0,
    ZKGPU_ERR_INVALID = -1,   /* bad argument (non power-of-two length, null pointer, cap too high ...) */
    ZKGPU_ERR_CUDA = -2,      /* CUDA runtime failure or no device */
    ZKGPU_ERR_ABORTED = -3,   /* abort flag observed (prover.rs:346-354 check_abort_signal) */
    ZKGPU_ERR_PROOF = -4,     /* proving failed (e.g. zeta in the trace subgroup, PoW search exhausted) */
    ZKGPU_ERR_NOMEM = -5
} zkgpu_status;

/* ZKGPU_MEM_AUTO (segment calls only): each table's pointer is host or device memory on its own (unified addressing decides), e.g. a
 * Keccak trace finished on the device (zkgpu_keccak_generate_trace) next to host traces of the other tables */
enum { ZKGPU_MEM_HOST = 0, ZKGPU_MEM_DEVICE = 1, ZKGPU_MEM_AUTO = 2 };

/* Table ids == `Table` enum, evm_arithmetization/src/all_stark.rs:74-86 */
enum {
    ZKGPU_TABLE_ARITHMETIC = 0, ZKGPU_TABLE_BYTE_PACKING = 1, ZKGPU_TABLE_CPU = 2, ZKGPU_TABLE_KECCAK = 3,
    ZKGPU_TABLE_KECCAK_SPONGE = 4, ZKGPU_TABLE_LOGIC = 5, ZKGPU_TABLE_MEMORY = 6, ZKGPU_TABLE_MEM_BEFORE = 7,
    ZKGPU_TABLE_MEM_AFTER = 8, ZKGPU_NUM_TABLES = 9
};

typedef struct zkgpu_ctx zkgpu_ctx;
typedef struct zkgpu_batch zkgpu_batch;     /* device-resident PolynomialBatch */
typedef struct zkgpu_ctl zkgpu_ctl;         /* device-resident CtlData of one table */
typedef struct zkgpu_proof zkgpu_proof;     /* host-resident StarkProofWithMetadata */

/* StarkConfig (starky config.rs) + FriConfig; standard_fast_config = {100, 2, 1, 4, 16, 4, 5, 84},
 * TEST_STARK_CONFIG (evm_arithmetization/src/testing_utils.rs:41-52) = {1, 1, 1, 4, 1, 4, 5, 1}. */
typedef struct {
    uint32_t security_bits;
    uint32_t num_challenges;
    uint32_t rate_bits;
    uint32_t cap_height;
    uint32_t proof_of_work_bits;
    uint32_t fri_arity_bits;        /* FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits) */
    uint32_t fri_final_poly_bits;
    uint32_t num_query_rounds;
} zkgpu_stark_config;

/* Constants of the assembled EVM kernel that appear in CpuStark constraints (KERNEL.global_labels[...]:
 * cpu/control_flow.rs:38-44, cpu/syscalls_exceptions.rs:68-73).  They come from assembling the 160 asm files,
 * which the Rust host does; the device evaluator takes them as parameters. */
typedef struct {
    uint64_t halt_final;
    uint64_t init;              /* label "init" (segment start), NOT "main" */
    uint64_t syscall_jumptable;
    uint64_t exception_jumptable;
} zkgpu_kernel_labels;

const char* zkgpu_last_error(void);
const char* zkgpu_version(void);

int zkgpu_ctx_create(int device, zkgpu_ctx** out);
void zkgpu_ctx_destroy(zkgpu_ctx* ctx);
int zkgpu_ctx_sync(zkgpu_ctx* ctx);
/* the context's CUDA stream (a cudaStream_t), so that a caller's own device work — e.g. the NCCL collectives of a table-sharded
 * segment — can be ordered with the library's kernels on the device instead of through host synchronisation */
int zkgpu_ctx_stream(zkgpu_ctx* ctx, void** stream_out);
/* CUDA-event stopwatch on the context's stream: device milliseconds between start and stop */
int zkgpu_ctx_timer_start(zkgpu_ctx* ctx);
int zkgpu_ctx_timer_stop(zkgpu_ctx* ctx, float* ms);
/* device-memory high-water mark and kernel-launch counter (bench.py's gpu_launches) */
int zkgpu_ctx_stats(zkgpu_ctx* ctx, uint64_t* kernel_launches, uint64_t* bytes_in_use, uint64_t* bytes_peak);

/* In-library profiler for bench.py's roofline numbers: while on, every launch group of a kernel family is bracketed by CUDA
 * events on the context's stream (the stream the kernels run on) and accounted with its ALGORITHMIC bytes.
 * family: 0 leaf_hash (Poseidon sponge over LDE rows), 1 merkle inner levels, 2 NTT / LDE passes, 3 quotient evaluation,
 * 4 CTL + lookup auxiliary columns, 5 openings, 6 FRI combine / fold / layer leaves, 7 proof-of-work grind,
 * 8 device-side trace finishing (zkgpu_keccak_generate_trace, zkgpu_logic_generate_trace,
 * zkgpu_arithmetic_generate_range_checks, zkgpu_memory_finish_trace). */
enum { ZKGPU_KF_LEAF_HASH = 0, ZKGPU_KF_MERKLE_LEVELS, ZKGPU_KF_NTT, ZKGPU_KF_QUOTIENT, ZKGPU_KF_AUX, ZKGPU_KF_OPENINGS, ZKGPU_KF_FRI,
       ZKGPU_KF_POW, ZKGPU_KF_TRACE_GEN, ZKGPU_KF_COUNT };
int zkgpu_ctx_set_profiling(zkgpu_ctx* ctx, int on);   /* on: also resets the counters */
int zkgpu_ctx_kernel_stats(zkgpu_ctx* ctx, uint32_t family, uint64_t* launches, double* ms_total, double* algorithmic_bytes);

/* Stage spans, the counterpart of the reference's TimingTree (`timed!` sections of prover.rs:92-116,137-143,296-341 and of starky's
 * prove_with_commitment): while on, every stage boundary of the prover records a CUDA event on the context's stream (no host
 * synchronisation, unlike ZKGPU_TRACE=1).  report: synchronises the stream and writes one "name<TAB>milliseconds" line per span, in
 * order — per segment "trace upload", one line per table commitment, "ctl data", then per table "aux columns", "aux commit",
 * "quotient eval", "quotient intt", "quotient commit", "openings eval", "observe openings (host)", "fri combine + intt",
 * "fri commit phase", "pow", "queries" and the table's name for the remainder.  *len in: capacity of buf, out: bytes needed (with
 * the terminating NUL); buf may be NULL to query.  set_timing (on or off) clears the recorded spans. */
int zkgpu_ctx_set_timing(zkgpu_ctx* ctx, int on);
int zkgpu_ctx_timing_report(zkgpu_ctx* ctx, char* buf, size_t* len);

/* ---- pinned host memory (for hosts that do not link the CUDA runtime themselves) ------------------------------------------- */
/* Trace uploads run at PCIe speed and asynchronously only from page-locked memory.  register: page-lock an
 * existing allocation in place (e.g. the Vec<u64> of a PolynomialValues) until unregister; alloc / free: a page-locked buffer.
 * Process-wide (any context / device may use the memory). */
int zkgpu_host_register(const void* ptr, size_t bytes);
int zkgpu_host_unregister(const void* ptr);
int zkgpu_host_alloc(size_t bytes, void** out);
int zkgpu_host_free(void* ptr);

/* ---- S1: commitments ------------------------------------------------------------------------------------- */
/* PolynomialBatch::from_values(values, rate_bits, blinding=false, cap_height): per column ifft -> zero-pad ->
 * coset FFT (shift = MULTIPLICATIVE_GROUP_GENERATOR) -> bit-reversed rows -> Poseidon Merkle tree.
 * `cols` = ncols pointers to n elements each (mem_kind says where they live).  keep_values != 0 keeps the raw
 * trace values on the device (needed later by zkgpu_ctl_data / lookup columns).  rate_bits 0..4: the STARK tables use 1
 * (StarkConfig::standard_fast_config), plonky2's recursion circuits commit with 3; the table prover (S3) takes rate_bits = 1 only. */
int zkgpu_commit_values(zkgpu_ctx* ctx, const uint64_t* const* cols, size_t ncols, size_t n, uint32_t rate_bits,
                        uint32_t cap_height, int mem_kind, int keep_values, zkgpu_batch** out);
/* same, but the columns are one contiguous column-major block: col c at base + c*n */
int zkgpu_commit_values_contig(zkgpu_ctx* ctx, const uint64_t* base, size_t ncols, size_t n, uint32_t rate_bits,
                               uint32_t cap_height, int mem_kind, int keep_values, zkgpu_batch** out);
/* PolynomialBatch::from_coeffs */
int zkgpu_commit_coeffs(zkgpu_ctx* ctx, const uint64_t* const* coeff_cols, size_t ncols, size_t n,
                        uint32_t rate_bits, uint32_t cap_height, int mem_kind, zkgpu_batch** out);
void zkgpu_batch_free(zkgpu_batch* b);
int zkgpu_batch_dims(const zkgpu_batch* b, size_t* ncols, size_t* n, uint32_t* rate_bits, uint32_t* cap_height);
/* merkle_tree.cap: (1 << cap_height) * 4 u64 */
int zkgpu_batch_cap(const zkgpu_batch* b, uint64_t* out_cap);
/* Fill the fields of a host PolynomialBatch (all buffers caller-owned, any may be NULL to skip):
 *   coeffs  ncols*n               polynomials[c].coeffs, column-major
 *   leaves  (n<<rate_bits)*ncols  merkle_tree.leaves, row-major, row j = LDE row bitrev(j)
 *   digests 2*((n<<rate_bits) - (1<<cap_height))*4   merkle_tree.digests in plonky2's recursive layout */
int zkgpu_batch_export(const zkgpu_batch* b, uint64_t* coeffs, uint64_t* leaves, uint64_t* digests);

/* Table jobs of this context (zkgpu_table_job_begin) also evaluate, in their challenger-independent first half, every constraint of the
 * table on the LDE — the values do not depend on the alphas, only their combination does (starky ConstraintConsumer: acc = acc alpha +
 * c).  zkgpu_table_job_finish then forms the quotient values as a Horner combination of the stored columns instead of running the
 * evaluator.  Same proofs bit for bit; more memory traffic in total, but the evaluator leaves the serial part of a table-sharded
 * segment (the tables are finished one after the other, prover.rs:251-259) — use it there, not for one-device proving. */
int zkgpu_ctx_set_precompute_constraints(zkgpu_ctx* ctx, int on);

/* ---- S1 split over several devices (the table-sharded layout of one segment) --------------------------------------
 * PolynomialBatch::from_values (prover.rs:100-107) of ONE table computed by k devices.  The transforms are per column and the leaf
 * sponge absorbs a row's columns in order, so: every device takes a column slice through ifft + LDE (zkgpu_lde_slice), the slices
 * are exchanged between the devices (the caller's collective: NCCL all-gather over NVLink), every device hashes the rows of one of
 * k equal blocks of leaves and builds the Merkle levels under that block's cap entries (zkgpu_merkle_block), the packed digests are
 * exchanged, and the owner of the table assembles the batch (zkgpu_batch_assemble).  The result is bit-identical to
 * zkgpu_commit_values on one device.  All buffers here are DEVICE memory owned by the caller, column-major. */
/* values: ncols x n (mem_kind: host or device).  Writes coeffs_out (ncols x n) and lde_out (ncols x (n << rate_bits), bit-reversed
 * rows, the layout of a batch); values_out (nullable) receives a device copy of the values when they came from the host. */
int zkgpu_lde_slice(zkgpu_ctx* ctx, const uint64_t* values, int mem_kind, size_t ncols, size_t n, uint32_t rate_bits,
                    uint64_t* values_out, uint64_t* coeffs_out, uint64_t* lde_out);
/* number of u64 words of one block's packed digests: sum over the levels (leaf digests up to the cap level) of 4 * count / nblocks */
int zkgpu_merkle_block_words(size_t nleaves, uint32_t cap_height, uint32_t nblocks, size_t* words);
/* digests of the leaves [block * nleaves / nblocks, (block + 1) * nleaves / nblocks) of the column-major LDE `lde` (column c at
 * lde + c * stride) and every Merkle level above them down to the cap level, packed level after level.  nblocks must divide the
 * cap (1 << cap_height).  Only the rows of the block are read: a device that holds just its block (pitch `stride` = rows per block,
 * received by an all-to-all) passes lde = block buffer - block * rows_per_block elements. */
int zkgpu_merkle_block(zkgpu_ctx* ctx, const uint64_t* lde, size_t stride, size_t ncols, size_t nleaves, uint32_t cap_height,
                       uint32_t nblocks, uint32_t block, uint64_t* packed_out);
/* A batch over caller-owned device buffers (borrowed until zkgpu_batch_free; values may be NULL): coeffs ncols x n with column
 * pitch coeff_pitch, lde ncols x N with column pitch N; packed = nblocks packed digest blocks of `zkgpu_merkle_block_words` words
 * each, unpacked into the batch's own digest levels. */
int zkgpu_batch_assemble(zkgpu_ctx* ctx, const uint64_t* values, const uint64_t* coeffs, const uint64_t* lde, const uint64_t* packed,
                         uint32_t nblocks, size_t ncols, size_t n, uint32_t rate_bits, uint32_t cap_height, zkgpu_batch** out);

/* ---- S2: cross-table-lookup data of one table -------------------------------------------------------------- */
/* Shape of a table under the AllStark registry (all_stark.rs:153-172): trace width and the auxiliary-column layout
 * lookup columns ++ CTL helper columns ++ CTL Z columns for `num_challenges` challenges. */
int zkgpu_table_info(uint32_t table_id, uint32_t num_challenges, uint32_t* num_columns, uint32_t* num_lookup_columns,
                     uint32_t* num_ctl_helper_columns, uint32_t* num_ctl_zs);
/* The per-table slice of starky get_ctl_data (prover.rs:137-143): for every cross-table lookup this table takes part in
 * and every challenge, the helper columns h = sum filter/combine(columns) and the running sum Z (ctls.md:17-25).
 * `trace` must have been committed with keep_values.  beta_gamma = [beta_0, gamma_0, beta_1, gamma_1]: the
 * GrandProductChallengeSet drawn by the caller's transcript after observing all trace caps and the public values. */
int zkgpu_ctl_data(zkgpu_ctx* ctx, uint32_t table_id, const zkgpu_batch* trace, const uint64_t* beta_gamma,
                   uint32_t num_challenges, zkgpu_ctl** out);
void zkgpu_ctl_free(zkgpu_ctl* ctl);
/* (num_ctl_helper_columns + num_ctl_zs) * n values, column-major: CtlData.zs_columns[*].helper_columns then [*].z */
int zkgpu_ctl_export(const zkgpu_ctl* ctl, uint64_t* out_cols);

/* ---- S3: prove_single_table (prover.rs:301-341) == starky prove_with_commitment ------------------------------- */
/* challenger_state: the 12-word sponge state of the shared transcript, in: as left by the previous table (or by
 * get_ctl_data for the first one), out: after this table's query challenges; buffers are compacted on both sides
 * exactly like `challenger.compact()` at prover.rs:320.  forced_pow_witness: NULL = search the smallest valid
 * witness; otherwise use the given one (plonky2 picks an arbitrary valid witness with rayon find_any, so a reference
 * proof can only be reproduced bit for bit when its witness is forced).  abort_flag: polled between stages
 * (prover.rs:317,346-354) -> ZKGPU_ERR_ABORTED.  labels may be NULL for every table but Cpu. */
int zkgpu_prove_table(zkgpu_ctx* ctx, uint32_t table_id, const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config,
                      const zkgpu_batch* trace, const zkgpu_ctl* ctl, uint64_t challenger_state[12],
                      const uint64_t* forced_pow_witness, volatile const int* abort_flag, zkgpu_proof** out);
void zkgpu_proof_free(zkgpu_proof* proof);
/* StarkProofWithMetadata as little-endian u64 words (layout: zk_evm_b200/csrc/stark/proof.h).  *len_words in: capacity
 * of buf, out: words needed; buf may be NULL to query the size. */
int zkgpu_proof_serialize(const zkgpu_proof* proof, uint64_t* buf, size_t* len_words);

/* ---- the shared Fiat-Shamir transcript: plonky2 Challenger<F, PoseidonHash> (prover.rs:118-144, 320) --------------------------- */
/* Host-side and tiny (a Poseidon permutation every 8 elements); exposed so a host without plonky2 (the Python mirror, a C++
 * harness) can replay prove_with_traces' transcript.  A Rust host keeps using plonky2's own Challenger and only passes the
 * 12-word state across the boundary. */
typedef struct zkgpu_challenger zkgpu_challenger;
int zkgpu_challenger_new(zkgpu_challenger** out);
void zkgpu_challenger_free(zkgpu_challenger* ch);
int zkgpu_challenger_observe(zkgpu_challenger* ch, const uint64_t* elements, size_t n);        /* observe_elements */
int zkgpu_challenger_get_challenges(zkgpu_challenger* ch, uint64_t* out, size_t n);            /* get_n_challenges */
int zkgpu_challenger_compact(zkgpu_challenger* ch, uint64_t state_out[12]);                    /* compact() */
int zkgpu_challenger_set_state(zkgpu_challenger* ch, const uint64_t state[12]);
/* prover.rs:118-144 in one call: observe the 9 trace caps in Table order (trace_caps = 9 x (4 << cap_height) words; an optional
 * table with table_in_use[t] == 0 contributes a zero cap, prover.rs:120-123), observe the flattened public values
 * (get_challenges.rs:202-227), draw (beta, gamma) x num_challenges into beta_gamma, return the compacted state the first
 * prove_single_table starts from. */
int zkgpu_segment_challenges(const uint64_t* trace_caps, const uint8_t* table_in_use, uint32_t cap_height, const uint64_t* public_values,
                             size_t n_public_values, uint32_t num_challenges, uint64_t* beta_gamma, uint64_t challenger_state[12]);

/* ---- prove_single_table split where the shared transcript is first needed (table-sharded multi-GPU runs) -------------------- */
/* begin: auxiliary polynomials (logUp columns ++ CTL helpers ++ CTL Zs) and their commitment — depends only on the trace and the
 * CTL challenges; finish: everything that consumes the transcript (alphas, quotient, zeta, openings, FRI, PoW, queries).
 * zkgpu_prove_table == begin + finish.  trace and ctl must outlive the job. */
typedef struct zkgpu_table_job zkgpu_table_job;
int zkgpu_table_job_begin(zkgpu_ctx* ctx, uint32_t table_id, const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config,
                          const zkgpu_batch* trace, const zkgpu_ctl* ctl, volatile const int* abort_flag, zkgpu_table_job** out);
int zkgpu_table_job_aux_cap(const zkgpu_table_job* job, uint64_t* out_cap, size_t* len_words);
int zkgpu_table_job_finish(zkgpu_table_job* job, uint64_t challenger_state[12], const uint64_t* forced_pow_witness,
                           volatile const int* abort_flag, zkgpu_proof** out);
void zkgpu_table_job_free(zkgpu_table_job* job);

/* ---- top: prove_with_traces (prover.rs:72-194) on one device ------------------------------------------------------------------ */
/* traces[t]: the trace of table t as one contiguous column-major block (column c at cols + c*n), or cols == NULL for an optional
 * table that is not in use (table_in_use[t] == false: zero cap observed, proof None -> proofs_out[t] == NULL).
 * public_values: the elements observe_public_values feeds the challenger, in order (the host flattens PublicValues).
 * forced_pow_witnesses: NULL, or 9 witnesses (entries of unused tables ignored).  ctl_challenges_out (nullable):
 * [beta_0, gamma_0, beta_1, gamma_1].  trace_caps_out (nullable): 9 x (4 << cap_height) words (MemBefore / MemAfter caps
 * become PublicValues.mem_before / mem_after, prover.rs:261-271). */
typedef struct {
    const uint64_t* cols;
    size_t n;
} zkgpu_table_trace;
int zkgpu_prove_segment(zkgpu_ctx* ctx, const zkgpu_table_trace* traces /*[ZKGPU_NUM_TABLES]*/, int mem_kind, const uint64_t* public_values,
                        size_t n_public_values, const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config,
                        const uint64_t* forced_pow_witnesses, volatile const int* abort_flag, zkgpu_proof** proofs_out /*[ZKGPU_NUM_TABLES]*/,
                        uint64_t* ctl_challenges_out, uint64_t* trace_caps_out);

/* The same call in two halves, for a stream of segments (zero/src/prover.rs:224-236 feeds segments one after the other): queue the
 * uploads of segment s+1 BEFORE proving segment s and the whole H2D chain of s+1 runs under the proof of s (its buffers are
 * allocated in stream order at the point of the call).  The traces must stay valid until zkgpu_prove_segment_uploaded returns
 * (pinned host memory, or the copies are staged synchronously).  An upload is consumed by exactly one prove call and freed with
 * zkgpu_upload_free (also when it was never proved).  zkgpu_prove_segment == upload + prove_uploaded. */
typedef struct zkgpu_upload zkgpu_upload;
int zkgpu_segment_upload(zkgpu_ctx* ctx, const zkgpu_table_trace* traces /*[ZKGPU_NUM_TABLES]*/, int mem_kind, const zkgpu_stark_config* config,
                         zkgpu_upload** out);
int zkgpu_prove_segment_uploaded(zkgpu_ctx* ctx, zkgpu_upload* upload, const uint64_t* public_values, size_t n_public_values,
                                 const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config, const uint64_t* forced_pow_witnesses,
                                 volatile const int* abort_flag, zkgpu_proof** proofs_out /*[ZKGPU_NUM_TABLES]*/, uint64_t* ctl_challenges_out,
                                 uint64_t* trace_caps_out);
void zkgpu_upload_free(zkgpu_upload* upload);

/* ---- device-side trace finishing (the data-parallel tail of trace generation) --------------------------------------- */
/* KeccakStark::generate_trace (keccak/keccak_stark.rs:70-250, called from witness/traces.rs into_tables): `inputs` = num_perms x 25
 * words (lane y*5 + x of each permutation's input state), `timestamps` = one per permutation, both host memory (borrowed for the call;
 * pinned memory must stay valid until the context is synchronised).  The trace is built in device memory: 2431 columns x n rows,
 * column-major, n = max(24 * num_perms, min_rows).next_power_of_two(), rows past 24 * num_perms all zero — bit for bit what
 * trace_rows_to_poly_values(generate_trace_rows(..)) holds.  The result is ordered on the context's stream: hand zkgpu_dev_trace_ptr to
 * zkgpu_prove_segment / zkgpu_segment_upload (ZKGPU_MEM_DEVICE or ZKGPU_MEM_AUTO) or zkgpu_commit_values_contig of the SAME context, or
 * zkgpu_ctx_sync first.  The 2431-column trace (2.5 GB at 2^17 rows) never crosses PCIe: 208 bytes per permutation do. */
typedef struct zkgpu_dev_trace zkgpu_dev_trace;
int zkgpu_keccak_generate_trace(zkgpu_ctx* ctx, const uint64_t* inputs, const uint64_t* timestamps, size_t num_perms, size_t min_rows,
                                zkgpu_dev_trace** out);
/* LogicStark::generate_trace (logic.rs:165-237): `ops` = num_ops x 9 words (host): operator (0 AND, 1 OR, 2 XOR), input0 and input1 as 4
 * little-endian u64 limbs each (U256).  523 columns x n rows, n = max(num_ops, min_rows).next_power_of_two(), padding rows zero. */
int zkgpu_logic_generate_trace(zkgpu_ctx* ctx, const uint64_t* ops, size_t num_ops, size_t min_rows, zkgpu_dev_trace** out);
/* A host trace (ncols x n contiguous column-major words) moved to the device as a zkgpu_dev_trace, for the in-place finishing steps */
int zkgpu_dev_trace_upload(zkgpu_ctx* ctx, const uint64_t* cols, size_t ncols, size_t n, zkgpu_dev_trace** out);
/* ArithmeticStark::generate_range_checks (arithmetic/arithmetic_stark.rs:130-156) in place on a device-resident Arithmetic trace
 * (116 columns x n >= 2^16 rows, as generate_trace pads it): RANGE_COUNTER[i] = min(i, 2^16 - 1); RC_FREQUENCIES[x] = number of cells
 * of the 96 shared columns equal to x.  What the two columns held before is ignored.  ZKGPU_ERR_INVALID if a shared cell is >= 2^16
 * (the reference asserts). */
int zkgpu_arithmetic_generate_range_checks(zkgpu_ctx* ctx, zkgpu_dev_trace* trace);
/* The data-parallel tail of MemoryStark::generate_trace (memory/memory_stark.rs:407-462).  `ops` (host): 14 x n words, column-major —
 * filter, timestamp, is_read, addr_context, addr_segment, addr_virtual, value_limbs[0..8] of every row AFTER the host's sort, fill_gaps
 * and pad_memory_ops (:215-236; n a power of two); `stale_contexts`: the list insert_stale_contexts takes (:387-404).  The device
 * derives the other 16 columns: timestamp_inv (into_row :104-131), the first-change flags, range_check, preinitialized_segments[_aux],
 * initialize_aux (generate_first_change_flags_and_rc :134-213), stale_contexts, is_pruned, and counter, frequencies,
 * stale_context_frequencies, is_stale, maybe_in_mem_after, mem_after_filter (generate_trace_col_major :240-294).  30 columns x n.
 * ZKGPU_ERR_INVALID if a range-checked difference does not fit the table (the reference asserts). */
int zkgpu_memory_finish_trace(zkgpu_ctx* ctx, const uint64_t* ops, size_t n, const uint64_t* stale_contexts, size_t num_stale,
                              zkgpu_dev_trace** out);
int zkgpu_dev_trace_dims(const zkgpu_dev_trace* t, size_t* ncols, size_t* n);
const uint64_t* zkgpu_dev_trace_ptr(const zkgpu_dev_trace* t);             /* device address of column 0 (column c at + c*n) */
int zkgpu_dev_trace_export(const zkgpu_dev_trace* t, uint64_t* host_out);  /* ncols * n words (parity tests) */
void zkgpu_dev_trace_free(zkgpu_dev_trace* t);

/* ---- stage-by-stage parity hooks (tests) ---------------------------------------------------------------------- */
/* when on, proofs retain their auxiliary / quotient PolynomialBatch and the FRI input values */
int zkgpu_ctx_set_debug(zkgpu_ctx* ctx, int on);
/* which: 0 = auxiliary polynomials, 1 = quotient chunks; the handle is borrowed from the proof */
int zkgpu_proof_debug_batch(const zkgpu_proof* proof, int which, const zkgpu_batch** out);
/* values of the FRI input polynomial on the LDE coset, bit-reversed order, (c0, c1) pairs */
int zkgpu_proof_debug_fri_values(const zkgpu_proof* proof, uint64_t* out, size_t* len_words);

/* ---- fine-grained entry points (per-kernel parity tests, ncu) --------------------------------------------- */
/* In-place batched NTT over `ncols` host columns of length n (column-major contiguous).
 * inverse=0: out[i] = sum_j in[j] w^(ij) (plonky2 fft, natural order in and out); inverse=1: ifft.
 * coset_shift != 0 and != 1: coset_fft(c, s) / coset_ifft(v, s). */
int zkgpu_ntt(zkgpu_ctx* ctx, uint64_t* data, size_t ncols, size_t n, int inverse, uint64_t coset_shift);
/* Poseidon permutation on `count` 12-element states (host memory) */
int zkgpu_poseidon_permute(zkgpu_ctx* ctx, uint64_t* states, size_t count);
/* hash_or_noop over `nrows` rows of `width` elements given column-major (col c at data + c*nrows): out nrows*4 */
int zkgpu_poseidon_hash_rows(zkgpu_ctx* ctx, const uint64_t* data_colmajor, size_t nrows, size_t width, uint64_t* out);
/* device-resident micro-benchmarks used by bench.py for the roofline numbers: run the kernel `iters` times on
 * synthetic device data and return the average milliseconds per launch measured with CUDA events on the ctx stream */
int zkgpu_bench_ntt(zkgpu_ctx* ctx, size_t ncols, size_t n, int iters, float* ms_per_iter, uint64_t* launches);
int zkgpu_bench_leaf_hash(zkgpu_ctx* ctx, size_t ncols, size_t nrows, int iters, float* ms_per_iter);
int zkgpu_bench_merkle_levels(zkgpu_ctx* ctx, size_t nleaves, int iters, float* ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* ZKGPU_H */
