/* grample_b200 — C ABI of the B200-native Gibbs hot path.
 *
 * This is the drop-in boundary for the reference's `sampler` package (the part of
 * CraigKelly/grample that cmd/root.go drives).  The reference is pure Go with no FFI of its
 * own; every entry point below names the Go function(s) it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the cgo binding a maintainer adds on the Go side.
 *
 * Conventions
 *   - opaque handles, plain pointers and sizes only (no C++/torch types);
 *   - every function returns 0 on success, non-zero on error; the message is available from
 *     gb_last_error() (thread-local), mirroring Go's `(value, error)` returns;
 *   - caller-allocated output buffers; host pointers unless a name says `_dev`;
 *   - every entry point selects its handle's CUDA device itself (goroutines migrate threads);
 *   - there is NO CPU fallback: without a CUDA device every compute call fails loudly;
 *   - thread safety: every entry point that takes a gb_chains* (or gb_fleet*) holds that handle's lock for the call,
 *     so goroutines / threads may call into one handle concurrently; the calls serialise in arrival order and the
 *     device work follows in stream order.  gb_chains_synchronize waits outside the lock.  A gb_model may be shared
 *     by handles driven from different threads (its lazily built tables are guarded).
 *
 * Layouts
 *   - a model is given as CSR arrays: card[n_vars], fixed[n_vars] (-1 = free), factor scopes
 *     scope_vars[scope_off[f] .. scope_off[f+1]) (first variable most significant, last
 *     fastest — model/function.go:180-202) and RAW (non-log) tables tables[tab_off[f] ..);
 *   - marginal vectors are flat: sum(card) doubles, variable-major;
 *   - chain state is uint8 `state[var][chain]` on the device (chain index fastest).
 */
#ifndef GRAMPLE_B200_H
#define GRAMPLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gb_model gb_model;   /* factor graph + sampler bookkeeping, host + device copies */
typedef struct gb_chains gb_chains; /* a population of chains (grouped by model) on ONE device   */
typedef struct gb_comm gb_comm;     /* one rank of an NCCL communicator, bound to one device      */
typedef struct gb_fleet gb_fleet;   /* the chain handles of one process's devices + their communicator */

/* sampler.NeighborVarMax (sampler/gibbs-collapsed.go:93) and model.maxTabSize (model/function.go:59) */
#define GB_NEIGHBOR_VAR_MAX 12
#define GB_MAX_TAB_SIZE (1 << 23)
#define GB_MAX_CARD 64 /* device limit on variable cardinality */

/* model.Measure implementations (model/error.go:81-249) */
enum gb_measure { GB_MAX_ABS = 0, GB_MEAN_ABS = 1, GB_HELLINGER = 2, GB_JS = 3 };
/* arithmetic of the sweep kernels.
 *   GB_F64   float64 log-sum-exp per update, 53-bit draws — follows the reference literally
 *   GB_F32   float32 log-sum-exp per update, 24-bit draws
 *   GB_TABLE conditionals evaluated in float64 ONCE per (variable, neighbour configuration) and
 *            stored as 32-bit inverse-CDF thresholds; the sweep is integer work (gather, index,
 *            compare) with 32-bit draws.  Applies when every sampled variable is binary with at
 *            most 256 joint configurations of its free neighbours (gb_model_table_mode). */
/*   GB_HYBRID per variable: variables of cardinality <= 4 with at most 65536 joint configurations of their free
 *            neighbours are sampled from threshold tables as in GB_TABLE (float64 conditional per
 *            configuration, stored as card - 1 cumulative 32-bit inverse-CDF thresholds; 32-bit draws, value =
 *            number of thresholds the draw exceeds), every other variable by the GB_F64 path (53-bit draws) —
 *            reference float64 arithmetic throughout, for models GB_TABLE rejects (collapsed variants
 *            with wide blankets, ternary / quaternary variables).  When every sampled variable has a table the whole
 *            model runs on the integer kernels.  Models with a cardinality above 4 run as GB_F64. */
/*   GB_TABLE_BITS the arithmetic and the law of GB_TABLE (float64 conditional per configuration, 32-bit thresholds,
 *            32-bit draws) on chain state packed one bit per chain, updated 32 chains at a time with bit-sliced
 *            logic (csrc/bits.cuh).  Applies when every sampled variable is binary with at most 4 free neighbours,
 *            all binary (gb_model_bits_mode) — the Ising / Grids problems; no per-chain histories (no
 *            GB_CHAINS_HISTORY), chain shards start on multiples of 32.  Its Philox stream is organised in bit planes,
 *            so trajectories differ from GB_TABLE's for the same seed; both are bit-checked against the oracle. */
enum gb_precision { GB_F64 = 0, GB_F32 = 1, GB_TABLE = 2, GB_HYBRID = 3, GB_TABLE_BITS = 4 };
/* gb_chains_create flags */
#define GB_CHAINS_HISTORY 1u /* keep per-chain half-window histograms (needed by gb_chains_convergence*) */
#define GB_CHAINS_PER_COLOUR 2u /* always launch one kernel per colour (disables the shared-memory-resident multi-sweep kernels small models use; same results) */
/* Rao-Blackwell marginal estimator (NOT the reference's estimator, hence a flag): a recorded update adds
 * the conditional it sampled from, p_k = e[k] / sum(e) (gibbs-simple.go:239-258 weights), to every bin of the variable
 * instead of Marginal[value] += 1 (chain.go:235).  Same expectation, lower variance.  Bins are 64-bit fixed point in
 * units of 2^-24 (gb_chains_group_counts returns them raw); merged marginals, TotalSampleCount, histories and
 * convergence scores keep their meaning.  GB_F64 / GB_F32: the floored log-sum-exp weights; GB_TABLE / GB_HYBRID: the
 * same conditional read back from the thresholds, p_0 = T_0 / 2^32, p_k = (T_k - T_{k-1}) / 2^32 (within 2^-32 of
 * the float64 value).  Not under GB_TABLE_BITS, nor for GB_TABLE models too large for the resident table kernel. */
#define GB_CHAINS_RAO_BLACKWELL 4u

const char* gb_last_error(void);
int gb_version(void);
int gb_device_count(int* n_out);

/* ------------------------------------------------------------------ model
 * gb_model_create   replaces sampler.NewGibbsSimple's bookkeeping (sampler/gibbs-simple.go:25-115:
 *                   Function.UseLogSpace with the 1e-6 eps rule model/function.go:126-142,
 *                   var->factor lists in m.Funcs order, validation :92-99) plus
 *                   GibbsCollapsed.FunctionsChanged (sampler/gibbs-collapsed.go:44-78: neighbour sets),
 *                   and builds the device CSR + colour schedule.
 * gb_model_load_uai replaces model.NewModelFromFile + UAIReader.ReadModel/ApplyEvidence
 *                   (model/model.go:52-112, model/uai.go:53-249); evid_path may be NULL. */
int gb_model_create(int32_t n_vars, const int32_t* card, const int32_t* fixed, int32_t n_funcs,
                    const int32_t* scope_off, const int32_t* scope_vars, const int64_t* tab_off,
                    const double* tables_raw, int device, gb_model** out);
int gb_model_load_uai(const char* uai_path, const char* evid_path, int device, gb_model** out);
void gb_model_destroy(gb_model* m);

int gb_model_n_vars(const gb_model* m, int32_t* out);
int gb_model_n_funcs(const gb_model* m, int32_t* out);
int gb_model_total_card(const gb_model* m, int32_t* out);            /* sum(card) */
int gb_model_cards(const gb_model* m, int32_t* out /*[n_vars]*/);
int gb_model_fixed(const gb_model* m, int32_t* out /*[n_vars]*/);
int gb_model_collapsed(const gb_model* m, int32_t* out /*[n_vars]*/);
/* factor f: arity/scope/table (log space, as the sampler sees it) — test & Go-shim read-back */
int gb_model_func_arity(const gb_model* m, int32_t f, int32_t* out);
int gb_model_func_scope(const gb_model* m, int32_t f, int32_t* out /*[arity]*/);
int gb_model_func_table_size(const gb_model* m, int32_t f, int64_t* out);
int gb_model_func_log_table(const gb_model* m, int32_t f, double* out /*[size]*/);
/* (*GibbsCollapsed).BlanketSize / FunctionCount (sampler/gibbs-collapsed.go:81-88) */
int gb_model_blanket_size(const gb_model* m, int32_t var, int32_t* out);
int gb_model_function_count(const gb_model* m, int32_t var, int32_t* out);
/* colour-sorted sweep order of the free, non-collapsed variables: order[n_order],
 * colour_off[n_colours+1].  Pass NULL outputs to query sizes only. */
int gb_model_schedule(const gb_model* m, int32_t* n_order, int32_t* n_colours, int32_t* order, int32_t* colour_off);

/* whether GB_TABLE applies to this model, and the total number of tabulated configurations */
/* hybrid mode: mask_out[n_vars] = 1 where the variable is sampled from a threshold table under GB_HYBRID */
int gb_model_hybrid_mask(const gb_model* m, int32_t* mask_out);
int gb_model_table_mode(gb_model* m, int32_t* ok_out, int64_t* n_thresholds_out);
/* whether GB_TABLE_BITS applies to this model */
int gb_model_bits_mode(const gb_model* m, int32_t* ok_out);
/* the tabulated thresholds of one sampled variable (builds the tables on first use): card - 1 cumulative thresholds
 * per configuration, out[configuration * (card - 1) + j] = the largest 32-bit draw that still selects a value <= j, so
 * the value drawn is the number of thresholds the draw exceeds (binary: value 0 iff draw <= threshold[configuration]);
 * configuration = sum(state[nbr_i] * stride_i) over the variable's free neighbours in ascending id order.
 * Pass out = NULL to query n = configurations * (card - 1). */
int gb_model_thresholds(gb_model* m, int32_t var, int32_t* n_out, uint32_t* out);

/* (*GibbsCollapsed).Collapse (sampler/gibbs-collapsed.go:98-314) as a pure function: returns a NEW
 * model in which `var` is summed out of its blanket (K3 collapse_marginalise on the device).
 * var < 0: pick uniformly among free, un-collapsed variables with blanket <= GB_NEIGHBOR_VAR_MAX,
 * at most n_vars tries (lines 102-120), using `seed`.  Same error cases as the reference
 * (fixed / already collapsed / empty new scope / table > GB_MAX_TAB_SIZE).
 * collapsed_var_out / marginal_out[card] may be NULL. */
int gb_model_collapse(const gb_model* src, int32_t var, uint64_t seed, int32_t* collapsed_var_out,
                      double* marginal_out, gb_model** out);

/* K5 conditional_probe — the parity hook for sampler/gibbs-simple.go:171-258: for a
 * caller-supplied full state, the floored un-normalised weights e[k] exactly as SampleVar
 * leaves them before WeightedSample (float64 or float32 arithmetic).  n_states states are
 * evaluated in one launch: states[n_states][n_vars], vars[n_states], out[n_states][GB_MAX_CARD]. */
int gb_conditional(const gb_model* m, int precision, int32_t n_states, const int32_t* states,
                   const int32_t* vars, double* out);

/* (*GibbsSimple).Sample / SampleVar and (*GibbsCollapsed).Sample for ONE caller-held state — the single-step form of
 * the FullSampler interface (sampler/sampler.go:16-22, gibbs-simple.go:148-271, gibbs-collapsed.go:317-334) that the
 * reference's benchmarks drive (gibbs-simple_test.go:68-86).  var < 0: a variable drawn uniformly among those with
 * FixedVal < 0 (and, with exclude_collapsed, not Collapsed; UniformSampler.VarSample, sampler.go:135-174), else that
 * variable (SampleVar).  The conditional is evaluated on the device (K5, `precision` GB_F64 or GB_F32), the
 * inverse-CDF draw (sampler.go:107-123) uses the Philox stream keyed by (seed, step): the caller passes a step
 * counter it increments.  state_inout[n_vars] is updated in place like the reference's `s []int`; *var_out = the
 * variable sampled (-1 on failure, as the reference returns).  One kernel launch and two small copies per step: an
 * API-compatibility path, not a throughput path — chains advance through gb_chains_*. */
int gb_model_sample(const gb_model* m, int precision, int32_t var, int exclude_collapsed, uint64_t seed, uint64_t step,
                    int32_t* state_inout, int32_t* var_out);

/* ------------------------------------------------------------------ chains
 * gb_chains_create replaces the chain-construction loop cmd/root.go:381-430 +
 * sampler.NewChain (sampler/chain.go:151-175) for ALL chains of one device at once:
 * group g holds chains_per_model[g] chains over models[g] (one group per distinct
 * collapsed variant; `simple` uses a single group).  Chain ids are global:
 * first_chain_id + local index; the Philox stream is keyed by (seed, global chain id) so
 * results do not depend on how chains are sharded across devices (shard sizes must be
 * multiples of 8).  Initial state: FixedVal or a uniform draw (gibbs-simple.go:103-111). */
int gb_chains_create(int32_t n_groups, gb_model* const* models, const int32_t* chains_per_model,
                     uint64_t seed, uint64_t first_chain_id, int precision, uint32_t flags, int device,
                     gb_chains** out);
/* adaptive.go:130-154: append a group of new chains over a (collapsed) model */
int gb_chains_add_group(gb_chains* c, gb_model* model, int32_t n_chains, uint64_t first_chain_id);
void gb_chains_destroy(gb_chains* c);
int gb_chains_n_groups(const gb_chains* c, int32_t* out);
int gb_chains_n_chains(const gb_chains* c, int64_t* out);

/* `n_sweeps` systematic colour sweeps of every chain; one sweep updates every free,
 * un-collapsed variable once (= n_free reference steps, chain.go:221-246).  record != 0 adds
 * each draw to the marginal counts (chain.go:231-236) and TotalSampleCount. */
int gb_chains_sweep(gb_chains* c, int64_t n_sweeps, int record);
/* gb_chains_sweep bracketed by CUDA events on the handle's own stream (device time of exactly
 * these sweeps, in milliseconds); returns after they finish.  Measurement hook for bench.py. */
int gb_chains_sweep_timed(gb_chains* c, int64_t n_sweeps, int record, float* ms_out);
/* number of kernels launched on behalf of this handle so far */
int gb_chains_launch_count(const gb_chains* c, int64_t* out);
/* Parity mode with the reference's OWN schedule: n_steps single-variable updates per chain, each
 * on a variable drawn uniformly among the free, un-collapsed ones ((*GibbsSimple).Sample /
 * (*GibbsCollapsed).Sample + UniformSampler.VarSample, sampler/sampler.go:135-174), float64
 * arithmetic.  One thread per chain: meant for small-model parity runs, not for throughput.
 * No window histograms are kept in this mode. */
int gb_chains_scan(gb_chains* c, int64_t n_steps, int record);
/* burn-in (chain.go:167-172): un-recorded.  The reference counts single-variable steps;
 * callers convert with ceil(steps / n_free). */
int gb_chains_burnin(gb_chains* c, int64_t n_sweeps);
/* One reference "round" for all chains — (*Chain).AdvanceChain (chain.go:180-218) + the
 * WaitGroup barrier (cmd/root.go:475-479): cw+1 recorded sweeps, the last 2*(cw/2) of which
 * fill the first/second half-window histograms (buffer/circular.go: the ring holds 2*(cw/2) values and Add
 * overwrites the oldest, so FirstHalf/SecondHalf split the NEWEST 2*(cw/2) samples — for an odd cw the round's
 * first two samples of a variable stay outside the window, for an even cw the first one). */
int gb_chains_advance(gb_chains* c, int32_t cw);
/* sum of Chain.TotalSampleCount over this device's chains (cmd/root.go:488-491) */
int gb_chains_total_samples(const gb_chains* c, int64_t* out);
/* Per-group forms.  A drop-in shim maps ONE reference *Chain to ONE group of replica chains:
 * NewChain's burn-in (chain.go:167-172) -> gb_chains_group_sweep(record = 0); (*Chain).AdvanceChain
 * -> gb_chains_group_advance, which only ENQUEUES the work on the handle's stream (like the
 * goroutine the reference spawns, chain.go:197-215); the caller's wg.Wait() is
 * gb_chains_synchronize.  gb_chains_group_info returns the group's chain count, its
 * TotalSampleCount and its (borrowed) model. */
int gb_chains_group_sweep(gb_chains* c, int32_t group, int64_t n_sweeps, int record);
int gb_chains_group_advance(gb_chains* c, int32_t group, int32_t cw);
int gb_chains_group_info(gb_chains* c, int32_t group, int32_t* n_chains_out, int64_t* total_samples_out,
                         gb_model** model_out);
int gb_chains_synchronize(gb_chains* c);

/* sampler.MergeChains (sampler/chain.go:96-148) over this device's chains: for every variable
 * the sum over chains of Marginal (each chain starts at uniform 1/card, model/variable.go:45,
 * plus its counts); a variable collapsed in ANY group is reported with that group's local
 * marginal and collapsed_out[v] = 1.  out[sum(card)], collapsed_out[n_vars] (may be NULL).
 * `out` is host memory; when it is page-locked (cudaHostAlloc / cudaHostRegister) the device-to-host copy lands in it
 * directly, otherwise it is staged through a pinned buffer of the handle. */
int gb_chains_merged_marginals(gb_chains* c, double* out, int32_t* collapsed_out);
/* The same, split so that the interval path overlaps with the next round's sweeps (the reference's monitor only
 * READS the merged marginals, cmd/root.go:498-539): _begin snapshots the counts in stream order and runs the reduction
 * (integer sums over this device's groups, NCCL sum over the communicator's ranks when one is attached, conversion to
 * marginals, copy to `out`) on a side stream; sweeps enqueued afterwards run concurrently with it.  _end blocks until
 * `out` / `collapsed_out` of the matching _begin are complete and returns the chain count and TotalSampleCount summed
 * over all ranks (either may be NULL).  Up to two merges may be pending per handle (begin, begin, end, begin, end, ...:
 * the host enqueues the next round and its snapshot before it waits for the previous result, so the device never runs
 * dry); _end completes the oldest.  `out` must stay valid until its _end; out == NULL takes part in the reduction
 * without a host copy of the marginals (ranks other than the reporting one).
 * The counts travel as 64-bit integers and the chains' uniform start mass is added once, after the reduction, so the
 * result is bit-identical however the chains are sharded over devices. */
int gb_chains_merge_begin(gb_chains* c, double* out, int32_t* collapsed_out);
int gb_chains_merge_end(gb_chains* c, int64_t* total_chains_out, int64_t* total_samples_out);
/* device time between the stage marks of the last completed merge, in milliseconds (measurement hook): ms_out[0] = count
 * sums on the sweep stream, [1] = NCCL sum over the ranks (includes waiting for SMs — the NCCL kernel starts when a
 * device-filling sweep kernel ends — and for the slowest rank), [2] = conversion to marginals (enqueued on the sweep stream
 * behind the next sweep kernel, so it includes that kernel when sweeps follow the merge), [3] = copy to the host */
int gb_chains_merge_timing(gb_chains* c, float* ms_out /*[4]*/);
/* chain count / TotalSampleCount over all ranks as of the last completed merge (cmd/root.go:488-491) */
int gb_chains_global_totals(const gb_chains* c, int64_t* total_chains_out, int64_t* total_samples_out);
/* Legacy multi-device form (the CALLER all-reduces; do not mix with an attached communicator): this device's un-merged contribution is written to a DEVICE buffer of
 * sum(card) doubles (collapsed variables zero) so the host plumbing can all-reduce it in
 * place (NCCL); finalize overwrites collapsed variables and copies to the host. */
int gb_chains_merge_partial_dev(gb_chains* c, double** dev_ptr_out, int64_t* n_out);
int gb_chains_merge_finalize(gb_chains* c, double* out, int32_t* collapsed_out);

/* sampler.ChainConvergence + (*Chain).ChainDist (chain.go:32-92, 253-290) with `measure`
 * (K4 on the device).  merged == NULL: merge this device's chains first.  out[n_vars]. */
int gb_chains_convergence(gb_chains* c, int measure, const double* merged, double* out);
/* Legacy multi-device form (the CALLER all-reduces): per-variable sums of within/between distances over this device's chains
 * into a DEVICE buffer [2*n_vars] (W then B) to be all-reduced, then finalised with the global
 * chain count. */
int gb_chains_convergence_partial_dev(gb_chains* c, int measure, const double* merged, double** dev_ptr_out,
                                      int64_t* n_out);
int gb_convergence_finalize(const gb_model* base, const double* wb /*[2*n_vars] host*/, int32_t cw,
                            int64_t total_chains, const int32_t* collapsed /*[n_vars]*/, double* out);

/* (*ConvergenceSampler).Adapt (sampler/adaptive.go:57-157): choose up to new_chain_count
 * variables (candidate filter :81-87, LOWEST convergence scores :102-119, ties by variable id),
 * collapse each on a fresh clone of `base`, append one group of chains_per_new_model chains per
 * variable with 2 burn-in steps (:145).  No-op at max_groups (reference MaxChains=128, :49).
 * chosen_out[new_chain_count] / n_chosen_out may be NULL.  The new models are owned by `c`. */
int gb_chains_adapt(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t chains_per_new_model,
                    int measure, int32_t cw, int32_t max_groups, uint64_t first_chain_id, int32_t* chosen_out,
                    int32_t* n_chosen_out);

/* With a communicator attached gb_chains_adapt is collective: the scores come from ChainConvergence over the chains of
 * every rank (identical on every rank, so all ranks collapse the same variables), chains_per_new_model counts a new
 * variant's chains over ALL ranks (this rank creates its block-aligned shard) and first_chain_id is the global id of the
 * first new variant's first chain; consecutive variants are ceil8(chains_per_new_model) ids apart.
 *
 * Explicit form of the same (caller-computed scores): `scores` [n_vars] are the ChainConvergence scores finalised from the
 * all-reduced within/between sums (gb_chains_convergence_partial_dev -> all-reduce ->
 * gb_convergence_finalize), identical on every rank, so every rank picks the same variables;
 * total_chains = chains over all ranks; chains_per_new_model = THIS rank's share of each new
 * variant's chains, first_chain_id = global id of this rank's first chain of the first new variant,
 * id_stride = global chains per variant (distance between consecutive variants' ids; 0 = padded
 * chains_per_new_model). */
int gb_chains_adapt_scores(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t chains_per_new_model,
                           const double* scores, int64_t total_chains, int32_t max_groups, uint64_t first_chain_id,
                           uint64_t id_stride, int32_t* chosen_out, int32_t* n_chosen_out);

/* state access for tests / the Go shim's LastSample: state[chain][var] int32 (host) */
int gb_chains_get_state(gb_chains* c, int32_t group, int32_t* out);
int gb_chains_set_state(gb_chains* c, int32_t group, const int32_t* in);
/* raw marginal counts of one group: uint64 [sum(card)] */
int gb_chains_group_counts(gb_chains* c, int32_t group, uint64_t* out);
/* per-chain half-window histograms of one group: uint16 [2][sum(card)][n_chains] */
int gb_chains_group_history(gb_chains* c, int32_t group, uint16_t* out);

/* ------------------------------------------------------------------ multi-GPU
 * The reference's only synchronisation is the round barrier (cmd/root.go:475-479) followed by MergeChains /
 * ChainConvergence on the host (:498-539, 640-668).  Here chains shard over devices by global chain id and those two
 * reductions run inside the library over NCCL (NVLink / NVSwitch): 64-bit integer counts (sum(card) + 2 words) and the
 * within / between sums (2 * n_vars + 1 doubles).  libnccl.so.2 is bound at run time (dlopen); single-device use
 * never touches it.
 *
 * One process per GPU: rank 0 calls gb_comm_unique_id and hands the 128 bytes to the other ranks by any means (the
 * launcher's store, a file, MPI); every rank calls gb_comm_init_rank with its device and attaches the communicator to
 * its chain handle.  From then on gb_chains_merged_marginals / gb_chains_merge_begin / gb_chains_convergence /
 * gb_chains_adapt are COLLECTIVE: every rank calls them in the same order. */
#define GB_COMM_ID_BYTES 128
int gb_comm_unique_id(uint8_t* id_out /*[GB_COMM_ID_BYTES]*/);
int gb_comm_init_rank(const uint8_t* id /*[GB_COMM_ID_BYTES], may be NULL when world == 1*/, int32_t world, int32_t rank,
                      int device, gb_comm** out);
int gb_comm_info(const gb_comm* comm, int32_t* world_out, int32_t* rank_out, int* device_out);
void gb_comm_destroy(gb_comm* comm);
/* the shard of `total_chains` chains that rank `rank` of `world` holds: contiguous, first id a multiple of 8 (chains share
 * Philox calls in blocks of 8), covering [0, total) over the ranks — the rule gb_chains_adapt / gb_fleet_adapt apply to a
 * new variant's chains, exposed so that a host shards its base chains the same way (pure host arithmetic) */
int gb_shard(int64_t total_chains, int32_t world, int32_t rank, uint64_t* first_out, int32_t* n_out);
/* comm may be NULL (detach).  The handle borrows the communicator: destroy the chains first. */
int gb_chains_attach_comm(gb_chains* c, gb_comm* comm);
/* One process, several GPUs (a Go or C++ host like cmd/root.go): a fleet owns a single-process communicator over
 * devices[] (ncclCommInitAll); the caller creates one model copy and one gb_chains per device (first_chain_id = the
 * shard's first global chain id) and attaches them.  The gb_fleet_* calls are the gb_chains_* calls of the same name
 * applied to every device — phase by phase, the NCCL calls of a phase inside one group, so ONE thread drives all
 * devices; results are written once (they are identical on every device).  gb_fleet_adapt takes one base model per
 * device slot (bases[i] lives on devices[i]). */
int gb_fleet_create(int32_t n_dev, const int* devices, gb_fleet** out);
void gb_fleet_destroy(gb_fleet* f);
int gb_fleet_size(const gb_fleet* f, int32_t* n_out);
int gb_fleet_attach(gb_fleet* f, int32_t slot, gb_chains* c);
int gb_fleet_sweep(gb_fleet* f, int64_t n_sweeps, int record);
int gb_fleet_advance(gb_fleet* f, int32_t cw);
int gb_fleet_synchronize(gb_fleet* f);
int gb_fleet_merged_marginals(gb_fleet* f, double* out, int32_t* collapsed_out);
int gb_fleet_merge_begin(gb_fleet* f, double* out, int32_t* collapsed_out);
int gb_fleet_merge_end(gb_fleet* f, int64_t* total_chains_out, int64_t* total_samples_out);
int gb_fleet_convergence(gb_fleet* f, int measure, const double* merged, double* out);
int gb_fleet_adapt(gb_fleet* f, gb_model* const* bases, int32_t new_chain_count, int32_t chains_per_new_model, int measure,
                   int32_t cw, int32_t max_groups, uint64_t first_chain_id, int32_t* chosen_out, int32_t* n_chosen_out);

/* ------------------------------------------------------------------ scoring (host)
 * model.NewErrorSuite (model/error.go:28-78).  fixed arrays may be NULL (= all free).
 * out8: MeanMeanAbs, MaxMeanAbs, MeanMaxAbs, MaxMaxAbs, MeanHellinger, MaxHellinger, MeanJS, MaxJS */
int gb_error_suite(int32_t n_vars, const int32_t* card, const int32_t* fixed1, const double* marg1,
                   const int32_t* fixed2, const double* marg2, double* out8);
/* UAIReader.ReadMargSolution (model/uai.go:252-332): card_out[n_vars], marg_out[sum(card)];
 * pass NULL outputs to query n_vars / total_card. */
int gb_mar_load(const char* path, int32_t* n_vars_out, int32_t* total_card_out, int32_t* card_out,
                double* marg_out);

#ifdef __cplusplus
}
#endif
#endif /* GRAMPLE_B200_H */
