// The compiled gradient tape: planner (fusion, memory, launch list) and executor.
//
// Replaces what `grad_compile(tape)` produces and `invokelatest(gf, f, args...; seed)` runs
// in the reference (/root/reference/src/grad.jl:293-307, 353-355): instead of one Julia
// statement -> one BLAS call / broadcast loop / mapreduce with a fresh temporary each, the op
// list is planned once into a short list of fused kernels over program-owned memory and
// replayed as a CUDA graph.
//
// Planning (host only, no device calls -- also runs on a machine without a GPU):
//   1. dead-code elimination from the OUTPUT tensors
//   2. fusion of elementwise ops (+ their sum / unbroadcast sinks and the tape's
//      broadcast(+, dx, old) accumulations) into regions: producer -> consumer and sibling
//      fusion whenever every other operand is already available when the region runs
//   3. region codegen: 2-d split (R x C), operand classes, shared-memory slot allocation,
//      micro-op list; lookup of a precompiled chain with the same micro-program
//   4. liveness-based arena planning for materialised intermediates and workspaces
#include <algorithm>
#include <cstring>
#include <chrono>
#include <string>

#include "common.h"

namespace yota {

static inline bool ew_is_binary(int op) { return op >= YOTA_EW_ADD && op <= YOTA_EW_DIV; }

enum StepKind {
  STEP_REGION,
  STEP_ROWSUM_FINISH,
  STEP_SCALAR_FINISH,
  STEP_COLSUM,
  STEP_GENERIC_EW,
  STEP_GENERIC_SUM,
  STEP_GEMM,
  STEP_GETINDEX,
  STEP_UNGETINDEX,
  STEP_LOGSOFTMAX,
  STEP_TRANSPOSE,
  STEP_CAT,
  STEP_SLICE,
  STEP_SLICE_ACCUM,  // out (aliasing acc) [I...] += dy : in-place form of ADD(UNGETINDEX(dy), acc)
  STEP_NOP           // a region that was absorbed as the epilogue of a GEMM
};

struct TInfo {
  yota_tensor_desc d;
  int nd = 0;             // canonical ndim (trailing 1s stripped)
  long long dims[4] = {1, 1, 1, 1};
  long long numel = 1;
  int producer = -1;
  std::vector<int> consumers;
  int alias_of = -1;  // tensor whose storage this one shares: RESHAPE outputs, contiguous getindex slices
  long long alias_elem_off = 0;  // ... starting this many elements into it
  bool inplace_alias = false;    // alias made by an in-place accumulation: the whole buffer, rewritten
  int alloc_step = -1;           // container of a zero-copy concatenation: first step that writes a member
  bool needed = false;
  bool materialized = false;
  int region = -1;  // region index if produced by a fused EW op
  // storage
  long long arena_off = -1;
  int def_step = -1, last_step = -1;
  int const_index = -1;
  int fwd_out_slot = -1;  // TEMP root whose in-place chain ends in an OUTPUT: lives in out_ptrs[slot]
};

struct Region {
  std::vector<int> ops;   // EW op indices, program order
  std::vector<int> sums;  // absorbed SUM op indices
  int nd = 0;
  long long dims[4] = {1, 1, 1, 1};
  unsigned kmask = 0;
  int step = -1;
  // codegen results
  int k = 0;
  RegionLaunch L;
  std::vector<int> gather_x, gather_idx;  // per RegionIn: gathered array / index vector (CLS_GATHER), else -1
  std::vector<int> in_tensors;   // tensor id per RegionIn
  std::vector<int> out_tensors;  // tensor id per RegionOut (for sums: the SUM's output tensor)
  std::vector<int> finish_steps;
  std::string sig;
  long long ws_off[RG_MAX_OUT];
  int n_partials[RG_MAX_OUT];
};

struct Step {
  int kind = 0;
  std::string label;
  std::vector<int> reads, writes;
  int region = -1;
  int op = -1;
  // finishers
  int out_tensor = -1;
  long long ws_off = -1;
  size_t ws_bytes = 0;
  long long R = 0, C = 0;
  int P = 0;
  float scale = 1.f;
  int ws_owner_step = -1;  // finisher: the step whose workspace it reads
  int shadow[2] = {-1, -1};  // GEMM (BF16 mode): index of the BF16 shadow of operand A / B
  bool cast[2] = {false, false};  // this step is the first user: cast before the GEMM
  int out_override = -1;  // in-place accumulation: the ADD's output tensor (aliases acc_tensor)
  int acc_tensor = -1;
  int epi_region = -1;    // GEMM: consumer region executed as the epilogue
  int epi_z_in = -1;      // index of the region input that is the GEMM output
  int tsh_make = -1;      // fused-epilogue GEMM: transposed shadow it also writes ...
  int tsh_make_out = -1;  // ... of this store output of the chain
  int tsh_use[2] = {-1, -1};  // GEMM: operand A / B is read from this transposed shadow (K-major)
  int rs_finish = -1;     // fused-epilogue GEMM: finisher step of its row-sum sink that runs inside the GEMM
  bool fused_away = false;  // finisher: performed by the last CTA of the kernel that wrote the partials
  int bsh_make = -1;      // fused-epilogue GEMM (BF16 mode): BF16 shadow it writes itself ...
  int bsh_make_out = -1;  // ... of this store output of the chain
  int fork = -1;          // small leaf step (plan_side_leaves): runs on a side stream, joined at the end of the sweep
  int unget_launches = 1; // kernels of an ungetindex step (sort passes + fold)
  int chain_len = 0;      // head of a recurrence (plan_recurrence_chains): this many GEMM + epilogue steps
  int chain_head = -1;    // member of a recurrence (the head too): index of its head step
  int chain_id = -1;      // head: which slice of the flag area is this recurrence's
  int ksplit = 0;         // GEMM cut along K inside the 2-CTA kernel; partial tiles in the step's workspace
  bool pdl = false;       // tcgen05 GEMM: launched with programmatic stream serialisation (plan_dependent_launch)
  bool early_a = false;   // ... and operand A is not written by the launch before it
  int presort = -1;       // scatter-add whose index vector is an input/constant: its sort runs on the side
                          // stream from the start of the run (index into presort_join), the step only folds
};

}  // namespace yota

using namespace yota;

struct yota_program {
  yota_ctx* ctx = nullptr;
  uint32_t flags = 0;
  std::vector<yota_op_desc> ops;
  std::vector<TInfo> T;
  std::vector<Region> regions;
  std::vector<Step> steps;
  std::vector<char> op_needed;
  // zero-copy concatenations made by rewrite_unrolled_loops: container tensor -> members, in order
  std::map<int, std::vector<int>> catv;
  // a concatenation that is a column range of an earlier one: (container tensor, element offset)
  std::map<int, std::pair<int, long long>> catv_view;
  int n_batched = 0;  // accumulation chains / slice families the rewrite collapsed into one contraction
  std::vector<int> op_scale;  // UNGETINDEX ops: scalar tensor multiplied into the source on load (-1: none)
  int n_in = 0, n_out = 0;
  size_t arena_bytes = 0, const_bytes = 0;
  size_t chain_flags_off = 0;  // step counters of the persistent recurrence launches: 256 bytes per recurrence
  int n_chains = 0;
  std::vector<int> chain_lens;    // steps of recurrence i ...
  std::vector<char> chain_taken;  // ... and whether the last run launched it persistently (before any run: assumed)
  size_t counter_off = 0, counter_bytes = 0;  // arrival counters of in-kernel reduction finishers (zeroed once)
  std::vector<int> const_tensors;  // CONST / SEED tensors, in const-buffer order
  int n_launches = 0, n_static = 0;
  // runtime
  char* arena = nullptr;
  std::vector<void*> bound_in, bound_out;
  float last_seed = 0.f;
  bool seed_valid = false;
  struct GraphEntry {
    std::vector<void*> in, out;
    cudaGraphExec_t exec = nullptr;
    int runs = 0;
    long long last_use = 0;
  };
  std::vector<GraphEntry> graphs;  // one captured graph per pointer binding (double-buffered batches)
  long long run_counter = 0;
  struct Shadow {
    int tensor = -1;  // root tensor id
    long long arena_off = -1;
    int first_step = -1, last_step = -1;
  };
  std::vector<Shadow> shadows;
  int shadow_elem_bytes = 2;  // BF16 operand shadows (2) or the Float32 lo shadows of the 3xTF32 mode (4)
  std::vector<Shadow> tshadows;  // transposed Float32 copies written by GEMM epilogues (TF32 mode)
  std::vector<int> allreduce_slots;
  std::vector<cudaEvent_t> ar_events;
  int n_presort = 0;
  int region_per_sm = 8;  // CTAs per SM the elementwise regions are sized for (fewer beside side-stream sorts)
  cudaEvent_t presort_fork = nullptr;
  std::vector<cudaEvent_t> presort_join;
  int n_forks = 0;
  std::vector<cudaEvent_t> fork_ev, join_ev;
  // where the caller-owned output buffers are touched, and which reductions follow which step
  // (plan_comm; rebuilt after yota_program_set_allreduce)
  struct CommPlan {
    bool valid = false;
    std::vector<long long> out_bytes, in_bytes;      // per out / in slot
    std::vector<std::vector<int>> first_touch_at;    // per step: out slots first touched there
    std::vector<std::vector<int>> ar_at;             // per step: all-reduced slots complete after it
    size_t last_ar_step = 0;
  } comm;
  // update!(x, gx) folded into the run (yota_program_set_update): param input slot, grad output slot
  std::vector<std::pair<int, int>> updates;
  float update_lr = 1.f;
  // Adam state of the marked pairs (yota_program_set_update_adam): per pair m and v, one step counter
  bool adam = false;
  float adam_b1 = 0.9f, adam_b2 = 0.999f, adam_eps = 1e-8f;
  std::vector<float*> adam_m, adam_v;
  int* adam_t = nullptr;
};

namespace yota {

static long long prod(const long long* d, int a, int b) {
  long long p = 1;
  for (int i = a; i < b; i++) p *= d[i];
  return p;
}

static bool is_ew(int opcode) { return opcode >= YOTA_EW_COPY && opcode <= YOTA_EW__LAST; }

// contraction mode of a program: 0 strict-FP32 FFMA, 1 TF32, 2 BF16, 3 compensated 3xTF32
static int gemm_mode_of(uint32_t flags) {
  return (flags & YOTA_FLAG_GEMM_TF32X3) ? 3 : (flags & YOTA_FLAG_GEMM_TF32) ? 1 : (flags & YOTA_FLAG_GEMM_BF16) ? 2 : 0;
}

static int root_of(const yota_program& P, int t) {
  while (P.T[t].alias_of >= 0) t = P.T[t].alias_of;
  return t;
}
// element offset of tensor t inside its root's storage
static long long alias_offset(const yota_program& P, int t) {
  long long off = 0;
  for (; P.T[t].alias_of >= 0; t = P.T[t].alias_of) off += P.T[t].alias_elem_off;
  return off;
}

// valid split points of an operand with canonical dims E (ne) inside iteration dims D (nd)
static unsigned operand_kmask(const long long* D, int nd, const long long* E, int ne) {
  unsigned m = 0;
  for (int k = 0; k <= nd; k++) {
    bool fullR = true, onesR = true, fullC = true, onesC = true;
    for (int i = 0; i < nd; i++) {
      const long long e = i < ne ? E[i] : 1;
      const bool full = e == D[i], one = e == 1;
      if (i < k) {
        fullR &= full;
        onesR &= one;
      } else {
        fullC &= full;
        onesC &= one;
      }
    }
    if ((fullR || onesR) && (fullC || onesC)) m |= 1u << k;
  }
  return m;
}

static int operand_class(const long long* D, int nd, int k, const long long* E, int ne) {
  bool fullR = true, fullC = true;
  for (int i = 0; i < nd; i++) {
    const long long e = i < ne ? E[i] : 1;
    if (i < k)
      fullR &= (e == D[i]);
    else
      fullC &= (e == D[i]);
  }
  if (fullR && fullC) return CLS_FULL;
  if (fullR) return CLS_ROWVEC;
  if (fullC) return CLS_COLVEC;
  return CLS_SCALAR;
}

// classification of a reduction mask at split k: 0 none/invalid, 1 ROWSUM, 2 SCALAR, 3 COLSUM
static int sum_class(const long long* D, int nd, int k, unsigned mask) {
  bool redR = true, keepR = true, redC = true, keepC = true;
  for (int i = 0; i < nd; i++) {
    if (D[i] == 1) continue;
    const bool r = mask >> i & 1;
    if (i < k) {
      redR &= r;
      keepR &= !r;
    } else {
      redC &= r;
      keepC &= !r;
    }
  }
  if (redR && redC) return 2;
  if (keepR && redC) return 1;
  if (redR && keepC) return 3;
  return 0;
}

static unsigned sum_kmask_fusable(const long long* D, int nd, unsigned mask) {
  unsigned m = 0;
  for (int k = 0; k <= nd; k++) {
    const int c = sum_class(D, nd, k, mask);
    if (c == 1 || c == 2) m |= 1u << k;
  }
  return m;
}

static bool same_dims(const TInfo& a, const long long* D, int nd) {
  if (a.nd != nd) return false;
  for (int i = 0; i < nd; i++)
    if (a.dims[i] != D[i]) return false;
  return true;
}

static const char* opname(int opc) {
  switch (opc) {
    case YOTA_EW_COPY: return "copy";
    case YOTA_EW_ADD: return "add";
    case YOTA_EW_SUB: return "sub";
    case YOTA_EW_MUL: return "mul";
    case YOTA_EW_DIV: return "div";
    case YOTA_EW_NEG: return "neg";
    case YOTA_EW_TANH: return "tanh";
    case YOTA_EW_EXP: return "exp";
    case YOTA_EW_LOG: return "log";
    case YOTA_EW_SIN: return "sin";
    case YOTA_EW_COS: return "cos";
    case YOTA_EW_SQRT: return "sqrt";
    case YOTA_EW_RELU: return "relu";
    case YOTA_EW_SIGMOID: return "sigmoid";
    case YOTA_EW_SQUARE: return "square";
    case YOTA_EW_GT0: return "gt0";
    case YOTA_EW_ADDC: return "addc";
    case YOTA_EW_MULC: return "mulc";
    case YOTA_EW_RSUBC: return "rsubc";
    case YOTA_EW_RDIVC: return "rdivc";
    case YOTA_EW_POWC: return "powc";
    case YOTA_OP_SUM: return "sum";
    case YOTA_OP_MATMUL: return "matmul";
    case YOTA_OP_GETINDEX: return "getindex";
    case YOTA_OP_UNGETINDEX: return "ungetindex";
    case YOTA_OP_LOGSOFTMAX: return "logsoftmax";
    case YOTA_OP_RESHAPE: return "reshape";
    case YOTA_OP_TRANSPOSE: return "transpose";
    case YOTA_OP_CAT: return "cat";
    case YOTA_OP_SLICE: return "slice";
    case 200: return "concat (members placed in one buffer)";
  }
  return "?";
}

// ------------------------------------------------------------------------------------------
// validation
// ------------------------------------------------------------------------------------------
static int validate(yota_program& P) {
  const int nt = (int)P.T.size();
  for (size_t i = 0; i < P.ops.size(); i++) {
    const yota_op_desc& o = P.ops[i];
    if (o.n_in < 0 || o.n_in > YOTA_MAX_OP_INPUTS) YOTA_FAIL(YOTA_E_ARG, "op %zu: bad n_in", i);
    if (o.out < 0 || o.out >= nt) YOTA_FAIL(YOTA_E_ARG, "op %zu: bad output tensor id", i);
    for (int j = 0; j < o.n_in; j++) {
      if (o.in[j] < 0 || o.in[j] >= nt) YOTA_FAIL(YOTA_E_ARG, "op %zu: bad input tensor id", i);
      const TInfo& t = P.T[o.in[j]];
      if (t.d.kind == YOTA_T_TEMP || t.d.kind == YOTA_T_OUTPUT)
        if (t.producer < 0 || t.producer >= (int)i)
          YOTA_FAIL(YOTA_E_ARG, "op %zu (%s): input %%%d is used before it is defined", i, opname(o.opcode), o.in[j]);
    }
    TInfo& out = P.T[o.out];
    if (out.d.kind != YOTA_T_TEMP && out.d.kind != YOTA_T_OUTPUT)
      YOTA_FAIL(YOTA_E_ARG, "op %zu writes a tensor that is not TEMP/OUTPUT", i);
    const int opc = o.opcode;
    if (is_ew(opc)) {
      const int want = ew_is_binary(opc) ? 2 : 1;
      if (o.n_in != want) YOTA_FAIL(YOTA_E_ARG, "op %zu (%s): expects %d inputs", i, opname(opc), want);
      for (int j = 0; j < o.n_in; j++) {
        const TInfo& a = P.T[o.in[j]];
        if (a.d.dtype != YOTA_F32) YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "op %zu (%s): only Float32 operands", i, opname(opc));
        for (int d = 0; d < std::max(a.nd, out.nd); d++) {
          const long long e = d < a.nd ? a.dims[d] : 1, D = d < out.nd ? out.dims[d] : 1;
          if (e != D && e != 1)
            YOTA_FAIL(YOTA_E_SHAPE, "op %zu (%s): DimensionMismatch, operand %%%d cannot broadcast to the output", i,
                      opname(opc), o.in[j]);
        }
      }
    } else if (opc == YOTA_OP_SUM) {
      const TInfo& x = P.T[o.in[0]];
      long long keep = 1;
      for (int d = 0; d < x.nd; d++)
        if (!(o.iattr[0] >> d & 1)) keep *= x.dims[d];
      if (keep != out.numel) YOTA_FAIL(YOTA_E_SHAPE, "op %zu (sum): output numel does not match the kept dims", i);
    } else if (opc == YOTA_OP_MATMUL) {
      const TInfo &A = P.T[o.in[0]], &B = P.T[o.in[1]];
      if (A.d.ndim > 2 || B.d.ndim > 2) YOTA_FAIL(YOTA_E_SHAPE, "op %zu (matmul): operands must be vectors/matrices", i);
      const long long ar = A.d.ndim >= 1 ? A.d.dims[0] : 1, ac = A.d.ndim == 2 ? A.d.dims[1] : 1;
      const long long br = B.d.ndim >= 1 ? B.d.dims[0] : 1, bc = B.d.ndim == 2 ? B.d.dims[1] : 1;
      const long long m = o.iattr[0] ? ac : ar, k1 = o.iattr[0] ? ar : ac;
      const long long k2 = o.iattr[1] ? bc : br, n = o.iattr[1] ? br : bc;
      if (k1 != k2 || m * n != out.numel)
        YOTA_FAIL(YOTA_E_SHAPE, "op %zu (matmul): DimensionMismatch (%lldx%lld)*(%lldx%lld)", i, m, k1, k2, n);
    } else if (opc == YOTA_OP_RESHAPE) {
      if (P.T[o.in[0]].numel != out.numel) YOTA_FAIL(YOTA_E_SHAPE, "op %zu (reshape): numel mismatch", i);
    } else if (opc == YOTA_OP_LOGSOFTMAX || opc == YOTA_OP_TRANSPOSE) {
      if (P.T[o.in[0]].numel != out.numel) YOTA_FAIL(YOTA_E_SHAPE, "op %zu (%s): numel mismatch", i, opname(opc));
    } else if (opc == 200 /* OP_CATV: made by rewrite_unrolled_loops, never by the caller */ && P.catv.count(o.out)) {
      long long cols = 0;
      for (int m : P.catv[o.out]) {
        if (m < 0 || m >= nt || P.T[m].producer < 0 || P.T[m].producer >= (int)i)
          YOTA_FAIL(YOTA_E_ARG, "internal: concatenation member %%%d is not defined before op %zu", m, i);
        cols += P.T[m].d.dims[1];
      }
      if (cols != out.d.dims[1]) YOTA_FAIL(YOTA_E_ARG, "internal: concatenation width mismatch");
    } else if (opc == YOTA_OP_GETINDEX || opc == YOTA_OP_UNGETINDEX || opc == YOTA_OP_CAT || opc == YOTA_OP_SLICE) {
      // checked when the step is built
    } else {
      YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "op %zu: unsupported opcode %d (no CPU fallback exists)", i, opc);
    }
  }
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// fusion
// ------------------------------------------------------------------------------------------
static int step_of_tensor(const yota_program& P, int t) {
  // step index at which tensor t becomes available (-1: before the program starts).  Aliases
  // carry their own step: a RESHAPE the step of its source, an in-place accumulation the step
  // that performs it (later than the step that produced the buffer it lives in).
  if (P.T[t].alias_of >= 0) return P.T[t].def_step;
  const TInfo& ti = P.T[root_of(P, t)];
  if (ti.d.kind == YOTA_T_INPUT || ti.d.kind == YOTA_T_CONST || ti.d.kind == YOTA_T_SEED) return -1;
  return ti.def_step;
}

static int new_step(yota_program& P, int kind, const std::string& label) {
  Step s;
  s.kind = kind;
  s.label = label;
  P.steps.push_back(s);
  return (int)P.steps.size() - 1;
}

// Elementwise placement.  T[t].region: -1 not placed, FLOATING waiting for a consumer,
// >= 0 member of that region.  An op whose operands come from no region of its shape (the
// broadcast-back of the seed, exp.(logp) of the logsoftmax pullback, ...) FLOATS until a
// consumer is placed and then joins the consumer's region, so it is never materialised.
constexpr int FLOATING = -2;

struct Fuser {
  yota_program& P;
  bool fuse;
  explicit Fuser(yota_program& p) : P(p), fuse(!(p.flags & YOTA_FLAG_NO_FUSE)) {}

  unsigned op_kmask(int opi) const {
    const yota_op_desc& o = P.ops[opi];
    const TInfo& out = P.T[o.out];
    unsigned km = ~0u;
    for (int j = 0; j < o.n_in; j++) km &= operand_kmask(out.dims, out.nd, P.T[o.in[j]].dims, P.T[o.in[j]].nd);
    return km & ((1u << (out.nd + 1)) - 1);
  }
  bool float_input(const yota_op_desc& o, int j) const { return P.T[o.in[j]].region == FLOATING; }

  // the floating producers feeding op `opi` (transitively), in program order, plus opi
  void group_of(int opi, std::vector<int>& g) const {
    const yota_op_desc& o = P.ops[opi];
    for (int j = 0; j < o.n_in; j++)
      if (float_input(o, j)) group_of(P.T[o.in[j]].producer, g);
    if (std::find(g.begin(), g.end(), opi) == g.end()) g.push_back(opi);
  }

  bool can_join(const std::vector<int>& g, int ri) const {
    const Region& r = P.regions[ri];
    unsigned km = r.kmask;
    if ((int)(r.ops.size() + g.size()) > 40) return false;
    std::vector<int> ins;
    auto in_group = [&](int t) {
      return P.T[t].producer >= 0 && std::find(g.begin(), g.end(), P.T[t].producer) != g.end();
    };
    for (int opi : r.ops)
      for (int j = 0; j < P.ops[opi].n_in; j++) {
        const int t = P.ops[opi].in[j];
        if (P.T[t].region != ri && std::find(ins.begin(), ins.end(), t) == ins.end()) ins.push_back(t);
      }
    for (int opi : g) {
      const yota_op_desc& o = P.ops[opi];
      if (!same_dims(P.T[o.out], r.dims, r.nd)) return false;
      km &= op_kmask(opi);
      for (int j = 0; j < o.n_in; j++) {
        const int t = o.in[j];
        if (P.T[t].region == ri || in_group(t)) continue;
        if (step_of_tensor(P, t) >= r.step) return false;  // operand not available when r runs
        if (std::find(ins.begin(), ins.end(), t) == ins.end()) ins.push_back(t);
      }
    }
    return km != 0 && (int)ins.size() <= RG_MAX_IN;
  }

  void do_join(const std::vector<int>& g, int ri) {
    Region& r = P.regions[ri];
    for (int opi : g) {
      r.kmask &= op_kmask(opi);
      r.ops.push_back(opi);
      TInfo& out = P.T[P.ops[opi].out];
      out.region = ri;
      out.def_step = r.step;
    }
  }

  int new_region(const TInfo& shape, unsigned km) {
    Region r;
    r.nd = shape.nd;
    for (int d = 0; d < 4; d++) r.dims[d] = shape.dims[d];
    r.kmask = km;
    r.step = new_step(P, STEP_REGION, "region");
    P.regions.push_back(r);
    const int ri = (int)P.regions.size() - 1;
    P.steps[r.step].region = ri;
    return ri;
  }

  // place op `opi` (and the floating producers it pulls along) now
  void place(int opi, bool allow_float) {
    const yota_op_desc& o = P.ops[opi];
    TInfo& out = P.T[o.out];
    if (!fuse) {
      do_join({opi}, new_region(out, op_kmask(opi)));
      return;
    }
    std::vector<int> g;
    group_of(opi, g);
    // candidate: the latest region that produces one of the group's operands of this shape
    int best = -1;
    for (int gi : g)
      for (int j = 0; j < P.ops[gi].n_in; j++) {
        const TInfo& a = P.T[P.ops[gi].in[j]];
        if (a.region >= 0 && same_dims(a, out.dims, out.nd) && a.region > best) best = a.region;
      }
    // Consumer-driven fusion: an op stays unplaced as long as nothing forces its value into
    // memory; it is placed together with the first consumer that must be (an output, the
    // operand of a GEMM / reduction / gather).  A value that is needed in memory anyway (the
    // layer activation H) is then never accompanied by stored derivatives of itself (1 - H^2 is
    // recomputed inside the pullback pass that consumes it).
    if (allow_float && group_inputs(g) <= RG_MAX_IN - 2 && g.size() <= 32) {
      out.region = FLOATING;
      return;
    }
    if (best >= 0 && can_join(g, best)) {
      do_join(g, best);
      return;
    }
    // sibling fusion: the latest region of this shape that already has every operand
    for (int ri = (int)P.regions.size() - 1; ri >= 0; ri--)
      if (!P.regions[ri].ops.empty() && can_join(g, ri)) {
        do_join(g, ri);
        return;
      }
    // a fresh region for the whole group, if it fits one
    unsigned km = ~0u;
    for (int gi : g) km &= op_kmask(gi);
    km &= (1u << (out.nd + 1)) - 1;
    if (km != 0 && g.size() <= 40 && group_inputs(g) <= RG_MAX_IN) {
      do_join(g, new_region(out, km));
      return;
    }
    // too big: settle the floating producers on their own first, then this op joins them
    bool any = false;
    for (int j = 0; j < o.n_in; j++)
      if (float_input(o, j)) {
        place(P.T[o.in[j]].producer, false);
        any = true;
      }
    if (any) {
      place(opi, false);
      return;
    }
    do_join({opi}, new_region(out, op_kmask(opi)));
  }

  int group_inputs(const std::vector<int>& g) const {
    std::vector<int> ins;
    for (int opi : g)
      for (int j = 0; j < P.ops[opi].n_in; j++) {
        const int t = P.ops[opi].in[j];
        const int pr = P.T[t].producer;
        if (pr >= 0 && std::find(g.begin(), g.end(), pr) != g.end()) continue;
        if (std::find(ins.begin(), ins.end(), t) == ins.end()) ins.push_back(t);
      }
    return (int)ins.size();
  }

  // a non-elementwise consumer (or the end of the program) needs tensor t in memory
  void force(int t) {
    if (P.T[t].region == FLOATING) place(P.T[t].producer, false);
  }
};

// Tape gradient accumulation  acc' = broadcast(+, new, acc)  (src/grad.jl:91) where `new` is a
// whole-array temporary with a single use: the reference allocates `new`, then a third array
// for the sum (C4: 127 x 3 passes over the 64 MiB dx, 127 x 2 over each dW).  When `acc` has
// no other reader either, the sum is formed IN PLACE in acc's buffer:
//   new = A*B (matmul rrule thunk)            ->  GEMM with C += ...  (same single add per element)
//   new = ungetindex of a slice (no index vec) ->  acc[I...] += dy; the other elements keep
//        their bits (the reference adds +0.0f to them: identical except -0.0f -> +0.0f)
static int count_needed_consumers(const yota_program& P, int t) {
  int n = 0;
  for (int c : P.T[t].consumers) n += P.op_needed[c] ? 1 : 0;
  return n;
}

static bool try_inplace_accumulate(yota_program& P, Fuser& F, int add_op) {
  const yota_op_desc& o = P.ops[add_op];
  TInfo& out = P.T[o.out];
  for (int side = 0; side < 2; side++) {
    const int a = o.in[side], b = o.in[1 - side];
    if (a == b) continue;
    const TInfo &ta = P.T[a], &tb = P.T[b];
    if (ta.producer < 0 || tb.producer < 0) continue;
    if (ta.d.kind != YOTA_T_TEMP || tb.d.kind != YOTA_T_TEMP) continue;
    if (ta.alias_of >= 0) continue;
    if (!same_dims(ta, out.dims, out.nd) || !same_dims(tb, out.dims, out.nd)) continue;
    if (count_needed_consumers(P, a) != 1 || count_needed_consumers(P, b) != 1) continue;
    // b's buffer must not be visible through another alias that is still read elsewhere: b is
    // its own root, or the end of a chain of earlier in-place accumulations (each link the whole
    // buffer, its only reader the add that replaced it).  A view of a larger or shared array
    // (x[:, :, t] of a saved activation, a reshape) is never accumulated into: the array's other
    // readers -- a later pullback region recomputing 1 - H^2 -- would see the sums.
    bool chain_ok = true;
    for (int t = b; P.T[t].alias_of >= 0; t = P.T[t].alias_of) chain_ok = chain_ok && P.T[t].inplace_alias;
    if (!chain_ok) continue;
    const int rb = root_of(P, b);
    if (P.T[rb].d.kind != YOTA_T_TEMP) continue;
    const yota_op_desc& pa = P.ops[ta.producer];
    int kind = -1;
    if (pa.opcode == YOTA_OP_MATMUL)
      kind = STEP_GEMM;
    else if (pa.opcode == YOTA_OP_UNGETINDEX) {
      bool vec = false;
      for (int d = 0; d < (int)pa.iattr[0]; d++) vec |= pa.iattr[1 + 3 * d] == YOTA_IDX_VEC;
      if (!vec) kind = STEP_SLICE_ACCUM;
    }
    if (kind < 0) continue;
    const int st = ta.def_step;
    if (st < 0 || P.steps[st].op != ta.producer || P.steps[st].out_override >= 0) continue;
    F.force(rb);
    if (step_of_tensor(P, b) >= st) continue;  // acc is not ready when the producer runs
    Step& step = P.steps[st];
    step.kind = kind;
    step.acc_tensor = b;
    step.out_override = o.out;
    step.reads.push_back(b);
    step.label += kind == STEP_GEMM ? " (+= in place)" : " -> slice += in place";
    P.T[rb].materialized = true;
    P.T[a].materialized = false;
    out.alias_of = b;
    out.inplace_alias = true;
    out.def_step = st;
    out.materialized = true;
    return true;
  }
  return false;
}

static int plan_fusion(yota_program& P) {
  Fuser F(P);
  const bool fuse = F.fuse;
  for (size_t i = 0; i < P.ops.size(); i++) {
    if (!P.op_needed[i]) continue;
    const yota_op_desc& o = P.ops[i];
    TInfo& out = P.T[o.out];
    if (!is_ew(o.opcode))
      for (int j = 0; j < o.n_in; j++) F.force(root_of(P, o.in[j]));
    if (o.opcode == 200 /* OP_CATV */) {
      // zero-copy concatenation: every member becomes a view of the container and is written
      // there by its own producer; the container is allocated when the first member is written
      if (P.catv_view.count(o.out)) {  // a column range of another concatenation: just an alias of it
        int last = -1;
        for (int m : P.catv[o.out]) {
          F.force(m);
          last = std::max(last, P.T[m].def_step);
        }
        out.alias_of = P.catv_view[o.out].first;
        out.alias_elem_off = P.catv_view[o.out].second;
        out.def_step = last;
        continue;
      }
      long long off = 0;
      int first = 1 << 30, last = -1;
      bool ok = true;
      for (int m : P.catv[o.out]) {
        F.force(m);
        TInfo& mt = P.T[m];
        ok = ok && mt.alias_of < 0 && mt.d.kind == YOTA_T_TEMP && mt.def_step >= 0;
        mt.alias_of = o.out;
        mt.alias_elem_off = off;
        mt.materialized = true;
        off += mt.numel;
        first = std::min(first, mt.def_step);
        last = std::max(last, mt.def_step);
      }
      if (!ok) YOTA_FAIL(YOTA_E_ARG, "internal: a concatenation member cannot be placed (alias / not a temporary)");
      out.def_step = last;
      out.alloc_step = first;
      out.materialized = true;
      continue;
    }
    if (o.opcode == YOTA_OP_RESHAPE) {
      out.alias_of = o.in[0];
      const int r = root_of(P, o.in[0]);
      P.T[r].materialized = true;
      out.def_step = step_of_tensor(P, o.in[0]);
      continue;
    }
    if (o.opcode == YOTA_OP_GETINDEX && fuse && out.d.kind == YOTA_T_TEMP && !getenv("YOTA_NO_SLICE_ALIAS")) {
      // x[:, ..., :, t] (or a range in the last place) of a column-major array is a contiguous
      // block: no copy, the result is a view (C4: the 128 x[:, :, t] slices of the unrolled RNN)
      const int n = (int)o.iattr[0];
      const TInfo& x = P.T[o.in[0]];
      bool contiguous = n >= 2 && n == x.d.ndim && x.d.dtype == YOTA_F32;
      for (int d = 0; d + 1 < n && contiguous; d++) contiguous = o.iattr[1 + 3 * d] == YOTA_IDX_COLON;
      const long long kind_last = contiguous ? o.iattr[1 + 3 * (n - 1)] : -1;
      contiguous = contiguous && (kind_last == YOTA_IDX_INT || kind_last == YOTA_IDX_RANGE);
      if (contiguous) {
        long long inner = 1;
        for (int d = 0; d + 1 < n; d++) inner *= x.d.dims[d];
        const long long a = o.iattr[2 + 3 * (n - 1)];
        const long long cnt = kind_last == YOTA_IDX_INT ? 1 : o.iattr[3 + 3 * (n - 1)] - a + 1;
        const long long off = (a - 1) * inner;
        if (a >= 1 && cnt >= 1 && a - 1 + cnt <= x.d.dims[n - 1] && inner * cnt == out.numel && off % 4 == 0) {
          out.alias_of = o.in[0];
          out.alias_elem_off = off;
          P.T[root_of(P, o.in[0])].materialized = true;
          out.def_step = step_of_tensor(P, o.in[0]);
          continue;
        }
      }
    }
    if (o.opcode == YOTA_EW_ADD && fuse && try_inplace_accumulate(P, F, (int)i)) continue;
    if (is_ew(o.opcode)) {
      if (F.op_kmask((int)i) == 0) {
        // broadcast pattern that does not collapse to 2-d: generic N-d kernel
        for (int j = 0; j < o.n_in; j++) F.force(root_of(P, o.in[j]));
        const int s = new_step(P, STEP_GENERIC_EW, std::string("ew_generic:") + opname(o.opcode));
        P.steps[s].op = (int)i;
        out.def_step = s;
        out.materialized = true;
        for (int j = 0; j < o.n_in; j++) P.T[root_of(P, o.in[j])].materialized = true;
        continue;
      }
      // operands of another shape (bias vectors, scalars) must be in memory
      for (int j = 0; j < o.n_in; j++)
        if (!same_dims(P.T[o.in[j]], out.dims, out.nd)) F.force(root_of(P, o.in[j]));
      F.place((int)i, out.d.kind != YOTA_T_OUTPUT);
      continue;
    }
    if (o.opcode == YOTA_OP_SUM) {
      const int x = o.in[0];
      const TInfo& xi = P.T[x];
      const unsigned mask = (unsigned)o.iattr[0];
      const unsigned skm = sum_kmask_fusable(xi.dims, xi.nd, mask);
      int target = -1;
      if (fuse && xi.region >= 0 && (P.regions[xi.region].kmask & skm) &&
          (int)P.regions[xi.region].sums.size() < 3)
        target = xi.region;
      if (target < 0 && skm != 0) {
        // stand-alone fixed-order reduction = a region with no micro-ops
        target = F.new_region(xi, (1u << (xi.nd + 1)) - 1);
        P.T[root_of(P, x)].materialized = true;
      }
      if (target >= 0) {
        Region& r = P.regions[target];
        r.kmask &= skm;
        r.sums.push_back((int)i);
        // the finisher step is created at codegen; reserve the availability step now
        const int fs = new_step(P, 0, "finish");
        P.steps[fs].op = (int)i;
        P.steps[fs].region = target;
        r.finish_steps.push_back(fs);
        out.def_step = fs;
        out.materialized = true;
        continue;
      }
      // column sums (sum over the leading dims) or arbitrary masks
      int kc = -1;
      for (int k = 1; k <= xi.nd; k++)
        if (sum_class(xi.dims, xi.nd, k, mask) == 3) {
          kc = k;
          break;
        }
      const int s = new_step(P, kc >= 0 ? STEP_COLSUM : STEP_GENERIC_SUM, kc >= 0 ? "colsum" : "sum_generic");
      P.steps[s].op = (int)i;
      if (kc >= 0) {
        P.steps[s].R = prod(xi.dims, 0, kc);
        P.steps[s].C = prod(xi.dims, kc, xi.nd);
      }
      P.steps[s].scale = (float)o.fattr[0];
      P.T[root_of(P, x)].materialized = true;
      out.def_step = s;
      out.materialized = true;
      continue;
    }
    // everything else: one step per op, all operands materialised
    int kind = 0;
    switch (o.opcode) {
      case YOTA_OP_MATMUL: kind = STEP_GEMM; break;
      case YOTA_OP_GETINDEX: kind = STEP_GETINDEX; break;
      case YOTA_OP_UNGETINDEX: kind = STEP_UNGETINDEX; break;
      case YOTA_OP_LOGSOFTMAX: kind = STEP_LOGSOFTMAX; break;
      case YOTA_OP_TRANSPOSE: kind = STEP_TRANSPOSE; break;
      case YOTA_OP_CAT: kind = STEP_CAT; break;
      case YOTA_OP_SLICE: kind = STEP_SLICE; break;
    }
    const int s = new_step(P, kind, opname(o.opcode));
    P.steps[s].op = (int)i;
    for (int j = 0; j < o.n_in; j++) P.T[root_of(P, o.in[j])].materialized = true;
    out.def_step = s;
    out.materialized = true;
  }
  // anything still floating has no placed consumer: place it now, in program order
  for (size_t i = 0; i < P.ops.size(); i++)
    if (P.op_needed[i] && is_ew(P.ops[i].opcode) && P.T[P.ops[i].out].region == FLOATING) F.place((int)i, false);
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// region codegen
// ------------------------------------------------------------------------------------------
static int codegen_region(yota_program& P, int ri) {
  Region& r = P.regions[ri];
  // split point
  int k = -1;
  for (int kk = 1; kk <= r.nd && k < 0; kk++)
    if ((r.kmask >> kk & 1) && prod(r.dims, 0, kk) % 4 == 0) k = kk;
  for (int kk = 1; kk <= r.nd && k < 0; kk++)
    if (r.kmask >> kk & 1) k = kk;
  if (k < 0) k = 0;
  if (!(r.kmask >> k & 1)) YOTA_FAIL(YOTA_E_ARG, "internal: region %d has no valid split", ri);
  r.k = k;
  RegionParams& p = r.L.p;
  memset(&p, 0, sizeof(p));
  p.R = prod(r.dims, 0, k);
  p.C = prod(r.dims, k, r.nd);
  r.L.V = (p.R % 4 == 0) ? 4 : 1;
  r.L.static_id = -1;

  // values: tensor id -> slot
  std::map<int, int> slot_of;
  std::vector<int> free_slots;
  int n_slots = 0;
  auto alloc_slot = [&]() {
    if (!free_slots.empty()) {
      const int s = free_slots.back();
      free_slots.pop_back();
      return s;
    }
    return n_slots++;
  };
  // inputs: FULL first (the kernel prefetches the first four FULL operands)
  std::vector<int> ins;
  auto add_in = [&](int t) {
    if (P.T[t].region == ri) return;
    if (std::find(ins.begin(), ins.end(), t) == ins.end()) ins.push_back(t);
  };
  for (int opi : r.ops)
    for (int j = 0; j < P.ops[opi].n_in; j++) add_in(P.ops[opi].in[j]);
  for (int si : r.sums) add_in(P.ops[si].in[0]);
  std::stable_sort(ins.begin(), ins.end(), [&](int a, int b) {
    const int ca = operand_class(r.dims, r.nd, k, P.T[a].dims, P.T[a].nd);
    const int cb = operand_class(r.dims, r.nd, k, P.T[b].dims, P.T[b].nd);
    return (ca == CLS_FULL) > (cb == CLS_FULL);
  });
  if ((int)ins.size() > RG_MAX_IN) YOTA_FAIL(YOTA_E_ARG, "internal: region %d has too many inputs", ri);
  p.n_in = (int)ins.size();
  r.in_tensors = ins;
  for (int q = 0; q < p.n_in; q++) {
    p.in[q].cls = operand_class(r.dims, r.nd, k, P.T[ins[q]].dims, P.T[ins[q]].nd);
    p.in[q].slot = alloc_slot();
    p.in[q].ptr = nullptr;
    slot_of[ins[q]] = p.in[q].slot;
    P.T[root_of(P, ins[q])].materialized = true;
  }
  // outputs: stores for values used outside the region, sinks for absorbed sums
  std::vector<int> store_t;
  for (int opi : r.ops) {
    const int t = P.ops[opi].out;
    bool ext = P.T[t].d.kind == YOTA_T_OUTPUT;
    for (int c : P.T[t].consumers) {
      if (!P.op_needed[c]) continue;
      const yota_op_desc& co = P.ops[c];
      const bool in_region_ew = is_ew(co.opcode) && P.T[co.out].region == ri;
      const bool in_region_sum =
          co.opcode == YOTA_OP_SUM && std::find(r.sums.begin(), r.sums.end(), c) != r.sums.end();
      if (!in_region_ew && !in_region_sum) ext = true;
    }
    if (ext) {
      store_t.push_back(t);
      P.T[t].materialized = true;
    }
  }
  if ((int)(store_t.size() + r.sums.size()) > RG_MAX_OUT)
    YOTA_FAIL(YOTA_E_ARG, "internal: region %d has too many outputs", ri);

  // last use (micro-op index) of every value; outputs/sinks are read after the op loop
  const int n_ops = (int)r.ops.size();
  std::map<int, int> last_use, n_uses;
  for (int q = 0; q < n_ops; q++)
    for (int j = 0; j < P.ops[r.ops[q]].n_in; j++) {
      last_use[P.ops[r.ops[q]].in[j]] = q;
      n_uses[P.ops[r.ops[q]].in[j]]++;
    }
  const int END = n_ops + 1;
  for (int t : store_t) last_use[t] = END, n_uses[t] += 2;
  for (int si : r.sums) last_use[P.ops[si].in[0]] = END, n_uses[P.ops[si].in[0]] += 2;

  p.n_ops = n_ops;
  if (n_ops > RG_MAX_OPS) YOTA_FAIL(YOTA_E_ARG, "internal: region %d has too many ops", ri);
  for (int q = 0; q < n_ops; q++) {
    const yota_op_desc& o = P.ops[r.ops[q]];
    MicroOp& m = p.ops[q];
    m.op = (uint8_t)o.opcode;
    m.imm = (float)o.fattr[0];
    const int prev_t = q > 0 ? P.ops[r.ops[q - 1]].out : -1;
    auto operand = [&](int t) -> uint8_t {
      if (t == prev_t) return RG_ACC;
      return (uint8_t)slot_of.at(t);
    };
    m.a = operand(o.in[0]);
    m.b = o.n_in > 1 ? operand(o.in[1]) : m.a;
    // free operand slots whose last use is this op (inputs keep theirs: reloaded per column)
    for (int j = 0; j < o.n_in; j++) {
      const int t = o.in[j];
      if (P.T[t].region == ri && last_use[t] == q && slot_of.count(t)) {
        free_slots.push_back(slot_of[t]);
        slot_of.erase(t);
      }
    }
    // destination: accumulator only if the single use is the very next micro-op
    const int t = o.out;
    const bool acc_only = n_uses[t] == 1 && last_use[t] == q + 1;
    if (acc_only || n_uses[t] == 0) {
      m.dst = RG_NONE;
    } else {
      const int s = alloc_slot();
      slot_of[t] = s;
      m.dst = (uint8_t)s;
    }
  }
  p.n_out = 0;
  r.out_tensors.clear();
  for (int t : store_t) {
    RegionOut& o = p.out[p.n_out++];
    o.kind = OUT_STORE;
    o.src = slot_of.at(t);
    o.acc = 0;
    o.ptr = nullptr;
    r.out_tensors.push_back(t);
  }
  for (size_t q = 0; q < r.sums.size(); q++) {
    const yota_op_desc& so = P.ops[r.sums[q]];
    RegionOut& o = p.out[p.n_out++];
    const int cls = sum_class(r.dims, r.nd, k, (unsigned)so.iattr[0]);
    o.kind = cls == 1 ? OUT_ROWSUM : OUT_SCALARSUM;
    o.src = slot_of.at(so.in[0]);
    o.acc = n_slots++;  // dedicated: accumulators persist across the column loop
    o.ptr = nullptr;
    r.out_tensors.push_back(so.out);
  }
  p.n_slots = n_slots;
  if (n_slots > RG_MAX_SLOTS) YOTA_FAIL(YOTA_E_ARG, "internal: region %d needs %d slots", ri, n_slots);
  region_geometry(r.L, P.ctx->sm_count, P.region_per_sm);
  r.sig = region_signature(p, r.L.V);
  if (!(P.flags & YOTA_FLAG_NO_STATIC)) r.L.static_id = static_chain_lookup(p, r.L.V);
  if (r.L.static_id >= 0) P.n_static++;

  // label + finisher steps
  std::string lab = "region[";
  for (int q = 0; q < n_ops; q++) lab += std::string(q ? "," : "") + opname(P.ops[r.ops[q]].opcode);
  lab += "]";
  for (size_t q = 0; q < r.sums.size(); q++) lab += "+sum";
  char buf[160];
  snprintf(buf, sizeof(buf), " %lldx%lld V%d in=%d out=%d slots=%d grid=%d%s%s", p.R, p.C, r.L.V, p.n_in, p.n_out,
           n_slots, r.L.grid, r.L.static_id >= 0 ? " static=" : " interp",
           r.L.static_id >= 0 ? static_chain_name(r.L.static_id) : "");
  Step& st = P.steps[r.step];
  st.label = lab + buf;
  for (int t : ins) st.reads.push_back(t);
  for (int t : store_t) st.writes.push_back(t);
  for (size_t q = 0; q < r.sums.size(); q++) {
    const int oi = (int)store_t.size() + (int)q;
    Step& fs = P.steps[r.finish_steps[q]];
    const yota_op_desc& so = P.ops[r.sums[q]];
    fs.out_tensor = so.out;
    fs.scale = (float)so.fattr[0];
    fs.region = ri;
    fs.ws_owner_step = r.step;
    fs.writes.push_back(so.out);
    if (p.out[oi].kind == OUT_ROWSUM) {
      fs.kind = STEP_ROWSUM_FINISH;
      fs.label = "rowsum_finish";
      fs.R = p.R;
      fs.P = p.colchunks * p.TCC;
      fs.ws_bytes = (size_t)fs.P * p.R * sizeof(float);
    } else {
      fs.kind = STEP_SCALAR_FINISH;
      fs.label = "scalar_finish";
      fs.P = r.L.grid;
      fs.ws_bytes = (size_t)r.L.grid * sizeof(float);
    }
    r.n_partials[oi] = fs.P;
  }
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// memory planning
// ------------------------------------------------------------------------------------------
struct FreeList {
  std::vector<std::pair<size_t, size_t>> blocks;  // (offset, size), sorted by offset
  size_t end = 0;
  size_t alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    int best = -1;
    for (size_t i = 0; i < blocks.size(); i++)
      if (blocks[i].second >= bytes && (best < 0 || blocks[i].second < blocks[best].second)) best = (int)i;
    if (best >= 0) {
      const size_t off = blocks[best].first;
      if (blocks[best].second == bytes)
        blocks.erase(blocks.begin() + best);
      else {
        blocks[best].first += bytes;
        blocks[best].second -= bytes;
      }
      return off;
    }
    // grow: extend a trailing free block if there is one
    if (!blocks.empty() && blocks.back().first + blocks.back().second == end) {
      const size_t off = blocks.back().first;
      end = off + bytes;
      blocks.pop_back();
      return off;
    }
    const size_t off = end;
    end += bytes;
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    auto it = std::lower_bound(blocks.begin(), blocks.end(), std::make_pair(off, (size_t)0));
    it = blocks.insert(it, {off, bytes});
    const size_t i = it - blocks.begin();
    if (i + 1 < blocks.size() && blocks[i].first + blocks[i].second == blocks[i + 1].first) {
      blocks[i].second += blocks[i + 1].second;
      blocks.erase(blocks.begin() + i + 1);
    }
    if (i > 0 && blocks[i - 1].first + blocks[i - 1].second == blocks[i].first) {
      blocks[i - 1].second += blocks[i].second;
      blocks.erase(blocks.begin() + i);
    }
  }
};

static size_t elem_size(const TInfo& t) { return t.d.dtype == YOTA_I64 ? 8 : 4; }

static void gemm_dims(const yota_program& P, const yota_op_desc& o, long long* M, long long* N, long long* K,
                      long long* lda, long long* ldb) {
  const TInfo &A = P.T[o.in[0]], &B = P.T[o.in[1]];
  const long long ar = A.d.ndim >= 1 ? A.d.dims[0] : 1, ac = A.d.ndim == 2 ? A.d.dims[1] : 1;
  const long long br = B.d.ndim >= 1 ? B.d.dims[0] : 1, bc = B.d.ndim == 2 ? B.d.dims[1] : 1;
  *M = o.iattr[0] ? ac : ar;
  *K = o.iattr[0] ? ar : ac;
  *N = o.iattr[1] ? br : bc;
  *lda = ar;
  *ldb = br;
}

// BF16 mode: every tensor-core-eligible GEMM reads BF16 shadows of its Float32 operands.  A
// shadow is cast once (by the first GEMM step that needs it) and shared by all later GEMMs of
// the same tensor (H_l feeds the forward and the dW GEMM, ΔZ_l the dW and dH GEMMs, ...).
// 3xTF32 mode: the same machinery plans the Float32 lo shadows  x - tf32(x)  (4 bytes per element,
// made by launch_split_tf32 before the first GEMM that reads the tensor).
static int plan_bf16_shadows(yota_program& P) {
  const int mode = gemm_mode_of(P.flags);
  if (mode != 2 && mode != 3) return YOTA_OK;
  P.shadow_elem_bytes = mode == 2 ? 2 : 4;
  const bool in_epilogue = mode == 2 && !(P.flags & YOTA_FLAG_NO_FUSE) && !getenv("YOTA_NO_EPI_SHADOW");
  std::map<int, int> shadow_of;
  for (size_t s = 0; s < P.steps.size(); s++) {
    Step& st = P.steps[s];
    if (st.kind != STEP_GEMM) continue;
    const yota_op_desc& o = P.ops[st.op];
    GemmArgs g{};
    g.mode = mode;
    g.transA = (int)o.iattr[0];
    g.transB = (int)o.iattr[1];
    gemm_dims(P, o, &g.M, &g.N, &g.K, &g.lda, &g.ldb);
    if (!gemm_tc_eligible(g)) continue;
    for (int j = 0; j < 2; j++) {
      const int root = root_of(P, o.in[j]);
      if (mode == 3 && (P.T[root].numel & 3)) continue;  // (never: eligible dims are multiples of 8)
      auto it = shadow_of.find(root);
      if (it == shadow_of.end()) {
        yota_program::Shadow sh;
        sh.tensor = root;
        sh.first_step = sh.last_step = (int)s;
        st.cast[j] = true;
        // the tensor is the stored output of a fused GEMM epilogue: that epilogue writes the BF16
        // copy too (one more 2-byte store per element it already holds) and the cast kernel --
        // a 4-byte read and a 2-byte write per element -- disappears
        const TInfo& rt = P.T[root];
        const int d = rt.def_step;
        if (in_epilogue && rt.d.kind == YOTA_T_TEMP && rt.alias_of < 0 && d >= 0 && d < (int)s &&
            P.steps[d].kind == STEP_GEMM && P.steps[d].epi_region >= 0 && P.steps[d].bsh_make < 0) {
          const Region& r = P.regions[P.steps[d].epi_region];
          const int n_store = r.L.p.n_out - (int)r.sums.size();
          for (int q = 0; q < n_store; q++)
            if (r.out_tensors[q] == root) {
              P.steps[d].bsh_make = (int)P.shadows.size();
              P.steps[d].bsh_make_out = q;
              P.steps[d].label += " +bf16";
              sh.first_step = d;
              st.cast[j] = false;
            }
        }
        P.shadows.push_back(sh);
        it = shadow_of.insert({root, (int)P.shadows.size() - 1}).first;
      }
      st.shadow[j] = it->second;
      P.shadows[it->second].last_step = (int)s;
    }
    if (st.shadow[0] == st.shadow[1]) st.cast[1] = false;
  }
  return YOTA_OK;
}

// TF32 mode: the dW GEMM of a layer, dW = dZ * H', reads BOTH operands MN-major (the batch, its
// K, is the slow axis of dZ and of H) and runs the tensor pipe at 66 % where K-major operands
// reach 94 % (profiles/README.md).  When such an operand is the stored output of a fused GEMM
// epilogue (H: forward bias+tanh; dZ: the dH GEMM's tanh' epilogue), that epilogue also writes a
// transposed copy -- each thread holds 32 consecutive columns of its row, one 128-byte line of
// the transpose -- and the dW GEMM reads the copy K-major.  Costs one extra store pass hidden
// behind the mainloop and numel * 4 bytes of arena while the copy is live.
static int plan_transposed_shadows(yota_program& P) {
  if (!(P.flags & YOTA_FLAG_GEMM_TF32) || (P.flags & YOTA_FLAG_NO_FUSE) || getenv("YOTA_NO_TSHADOW")) return YOTA_OK;
  std::map<int, int> shadow_of;
  for (size_t s = 0; s < P.steps.size(); s++) {
    Step& st = P.steps[s];
    if (st.kind != STEP_GEMM || st.epi_region >= 0) continue;
    const yota_op_desc& o = P.ops[st.op];
    GemmArgs g{};
    g.mode = 1;
    g.transA = (int)o.iattr[0];
    g.transB = (int)o.iattr[1];
    gemm_dims(P, o, &g.M, &g.N, &g.K, &g.lda, &g.ldb);
    if (!gemm_tc2_eligible(g, P.ctx->sm_count)) continue;
    for (int j = 0; j < 2; j++) {
      const bool mn_major = j == 0 ? !g.transA : g.transB != 0;
      if (!mn_major) continue;
      const int t = o.in[j];
      const TInfo& ti = P.T[t];
      if (ti.alias_of >= 0 || ti.d.kind != YOTA_T_TEMP || !ti.materialized || ti.d.ndim != 2) continue;
      if (ti.dims[1] % 32 || ti.dims[0] % 4) continue;  // 128-byte lines of the transpose, TMA strides
      const int d = ti.def_step;
      if (d < 0 || d >= (int)s) continue;
      Step& ps = P.steps[d];
      if (ps.kind != STEP_GEMM || ps.epi_region < 0) continue;
      const Region& r = P.regions[ps.epi_region];
      const int n_store = r.L.p.n_out - (int)r.sums.size();
      int q_out = -1;
      for (int q = 0; q < n_store; q++)
        if (r.out_tensors[q] == t) q_out = q;
      if (q_out < 0) continue;
      auto it = shadow_of.find(t);
      if (it == shadow_of.end()) {
        if (ps.tsh_make >= 0) continue;  // one transposed copy per epilogue
        yota_program::Shadow sh;
        sh.tensor = t;
        sh.first_step = d;
        sh.last_step = (int)s;
        P.tshadows.push_back(sh);
        it = shadow_of.insert({t, (int)P.tshadows.size() - 1}).first;
        ps.tsh_make = it->second;
        ps.tsh_make_out = q_out;
        ps.label += " +T";
      }
      st.tsh_use[j] = it->second;
      P.tshadows[it->second].last_step = std::max(P.tshadows[it->second].last_step, (int)s);
    }
    if (st.tsh_use[0] >= 0 || st.tsh_use[1] >= 0)
      st.label += std::string(" [K-major ") + (st.tsh_use[0] >= 0 ? "A" : "") + (st.tsh_use[1] >= 0 ? "B" : "") + "]";
  }
  return YOTA_OK;
}

// tensors a step reads (as the liveness pass sees them) / writes
static std::vector<int> step_reads(const yota_program& P, const Step& st) {
  std::vector<int> r = st.reads;
  if (st.op >= 0 && st.kind != STEP_ROWSUM_FINISH && st.kind != STEP_SCALAR_FINISH)
    for (int j = 0; j < P.ops[st.op].n_in; j++) r.push_back(P.ops[st.op].in[j]);
  return r;
}
static std::vector<int> step_writes(const Step& st) {
  std::vector<int> w = st.writes;
  if (st.out_tensor >= 0) w.push_back(st.out_tensor);
  return w;
}

// A launch-bound program (C1: twelve kernels of a few microseconds, each waiting for the one before)
// is a chain only because a stream is one: dW, db and the loss are LEAVES of the dataflow -- nothing
// later reads or rewrites what they produce.  Small leaf steps (finishers, column sums, GEMMs below
// ~0.5 GFLOP) whose results are program outputs are forked onto side streams right where they stand
// and joined at the end of the sweep; inside the CUDA graph they are parallel branches, and the
// critical path is what really depends on what (C1: 12 -> 7 launches deep).  What a forked step
// reads stays allocated until the join (plan_memory).  Big kernels fill the GPU anyway and stay in
// line.  YOTA_NO_FORK=1 turns it off; programs that all-reduce are issued in stream order.
static int plan_side_leaves(yota_program& P) {
  if (getenv("YOTA_NO_FORK")) return YOTA_OK;
  const int ns = (int)P.steps.size();
  int last_launch = -1;
  for (int s = 0; s < ns; s++)
    if (P.steps[s].kind != STEP_NOP && !P.steps[s].fused_away) last_launch = s;
  for (int s = 0; s < ns; s++) {
    Step& st = P.steps[s];
    if (st.kind == STEP_NOP || st.fused_away || s == last_launch) continue;
    bool small = st.kind == STEP_ROWSUM_FINISH || st.kind == STEP_SCALAR_FINISH || st.kind == STEP_COLSUM;
    if (st.kind == STEP_GEMM && st.epi_region < 0 && st.chain_head < 0 && st.shadow[0] < 0 && st.shadow[1] < 0 &&
        st.tsh_use[0] < 0 && st.tsh_use[1] < 0) {
      long long M, N, K, lda, ldb;
      gemm_dims(P, P.ops[st.op], &M, &N, &K, &lda, &ldb);
      small = 2.0 * (double)M * (double)N * (double)K <= 5.4e8;
    }
    if (!small) continue;
    const std::vector<int> w = step_writes(st), r = step_reads(P, st);
    bool leaf = !w.empty();
    for (int t : w) leaf = leaf && P.T[root_of(P, t)].d.kind == YOTA_T_OUTPUT;  // a caller's buffer: never recycled
    for (int s2 = s + 1; s2 < ns && leaf; s2++) {
      const Step& o = P.steps[s2];
      if (o.kind == STEP_NOP) continue;
      for (int t2 : step_writes(o)) {  // nobody rewrites what it writes or reads
        for (int t : w) leaf = leaf && root_of(P, t2) != root_of(P, t);
        for (int t : r) leaf = leaf && root_of(P, t2) != root_of(P, t);
      }
      for (int t2 : step_reads(P, o))  // nobody reads what it writes
        for (int t : w) leaf = leaf && root_of(P, t2) != root_of(P, t);
    }
    if (!leaf) continue;
    st.fork = P.n_forks++;
    st.label += " (leaf: on a side stream)";
  }
  return YOTA_OK;
}

// Back-to-back skinny GEMMs (the per-step products of an unrolled recurrence: 1024 x 128 x 1024,
// ~9 us each, most of it launch latency, prologue and the first TMA round trip) are launched
// with programmatic stream serialisation: the next grid is scheduled while this one drains, and
// waits (griddepcontrol.wait) before it reads B, epilogue operands or writes anything.  Where the
// launch just before does not write this GEMM's A operand (the recurrence weights), the first A
// tiles are requested before that wait.  Only the 1-CTA kernel looks at these flags; every other
// kernel is launched and ordered as before.  YOTA_NO_PDL=1 turns it off.
// An unrolled recurrence  h_i = chain(W * h_{i-1}, ...)  shows up as a run of consecutive skinny
// GEMM + epilogue steps of one shape and one chain whose B operand is the previous step's stored
// output.  Such a run is marked here; at run time (try_run_chain) the bound addresses are checked
// for constant strides and the whole run becomes ONE persistent launch (tc::gemm_chain_kernel:
// weights resident in shared memory, steps ordered through per-tile flags) -- or, if anything
// does not fit, the steps are launched one by one as before.  YOTA_NO_CHAIN=1 turns it off.
static int plan_recurrence_chains(yota_program& P) {
  if (getenv("YOTA_NO_CHAIN") || gemm_mode_of(P.flags) != 1) return YOTA_OK;
  std::vector<int> L;
  for (size_t s = 0; s < P.steps.size(); s++)
    if (P.steps[s].kind != STEP_NOP && !P.steps[s].fused_away) L.push_back((int)s);
  auto cand = [&](int s) {
    const Step& st = P.steps[s];
    return st.kind == STEP_GEMM && st.epi_region >= 0 && st.tsh_make < 0 && st.bsh_make < 0 && st.rs_finish < 0 &&
           st.acc_tensor < 0 && st.shadow[0] < 0 && st.tsh_use[0] < 0 && st.tsh_use[1] < 0 &&
           P.regions[st.epi_region].sums.empty() && P.regions[st.epi_region].L.static_id >= 0;
  };
  auto same = [&](int a, int b) {
    const yota_op_desc &oa = P.ops[P.steps[a].op], &ob = P.ops[P.steps[b].op];
    long long da[5], db[5];
    gemm_dims(P, oa, &da[0], &da[1], &da[2], &da[3], &da[4]);
    gemm_dims(P, ob, &db[0], &db[1], &db[2], &db[3], &db[4]);
    for (int q = 0; q < 5; q++)
      if (da[q] != db[q]) return false;
    return oa.iattr[0] == ob.iattr[0] && oa.iattr[1] == ob.iattr[1] && P.steps[a].epi_z_in == P.steps[b].epi_z_in &&
           P.regions[P.steps[a].epi_region].L.static_id == P.regions[P.steps[b].epi_region].L.static_id &&
           root_of(P, oa.in[0]) == root_of(P, ob.in[0]) && alias_offset(P, oa.in[0]) == alias_offset(P, ob.in[0]);
  };
  auto feeds = [&](int a, int b) {  // B of step b is a stored output of step a's epilogue
    const int bt = P.ops[P.steps[b].op].in[1];
    for (int t : P.regions[P.steps[a].epi_region].out_tensors)
      if (t == bt || (root_of(P, t) == root_of(P, bt) && alias_offset(P, t) == alias_offset(P, bt) &&
                      P.T[t].numel == P.T[bt].numel))
        return true;
    return false;
  };
  for (size_t i = 0; i < L.size();) {
    if (!cand(L[i])) {
      i++;
      continue;
    }
    size_t j = i + 1;
    while (j < L.size() && cand(L[j]) && same(L[i], L[j]) && feeds(L[j - 1], L[j])) j++;
    const int n = (int)(j - i);
    if (n >= 4) {
      Step& head = P.steps[L[i]];
      head.chain_len = n;
      head.chain_id = P.n_chains++;
      P.chain_lens.push_back(n);
      P.chain_taken.push_back(1);
      head.label += " [recurrence of " + std::to_string(n) + " steps: one persistent launch]";
      for (size_t q = i; q < j; q++) P.steps[L[q]].chain_head = L[i];
    }
    i = j;
  }
  return YOTA_OK;
}

// Contractions with few output tiles and a long K (the batched dW = dZ_all * H_all' of an unrolled
// recurrence: 1024 x 1024 x 16384) are cut along K inside the 2-CTA kernel; the partial tiles need a
// workspace for the duration of the step.
static int plan_split_k(yota_program& P) {
  if (gemm_mode_of(P.flags) != 1) return YOTA_OK;
  for (Step& st : P.steps) {
    if (st.kind != STEP_GEMM || st.epi_region >= 0) continue;
    const yota_op_desc& o = P.ops[st.op];
    GemmArgs g{};
    g.mode = 1;
    g.transA = (int)o.iattr[0];
    g.transB = (int)o.iattr[1];
    gemm_dims(P, o, &g.M, &g.N, &g.K, &g.lda, &g.ldb);
    g.ldc = g.M;
    if (st.tsh_use[0] >= 0) g.lda = g.K;  // (read from a K-major copy: does not change eligibility)
    if (st.tsh_use[1] >= 0) g.ldb = g.K;
    const int ks = gemm_tc2_ksplit(g, P.ctx->sm_count);
    if (ks > 1) {
      st.ksplit = ks;
      st.ws_bytes = (size_t)ks * (size_t)g.M * (size_t)g.N * sizeof(float);
      st.label += " [K split in " + std::to_string(ks) + "]";
    }
  }
  return YOTA_OK;
}

static int plan_dependent_launch(yota_program& P) {
  if (getenv("YOTA_NO_PDL") || gemm_mode_of(P.flags) != 1) return YOTA_OK;
  int prev = -1;
  for (size_t s = 0; s < P.steps.size(); s++) {
    Step& st = P.steps[s];
    if (st.kind == STEP_NOP || st.fused_away || st.fork >= 0) continue;  // (forked leaves are not in this stream)
    if (st.kind == STEP_GEMM && prev >= 0) {
      st.pdl = true;
      const int a_root = root_of(P, P.ops[st.op].in[0]);
      bool clean = st.tsh_use[0] < 0 || P.tshadows[st.tsh_use[0]].first_step != prev;
      for (int t : P.steps[prev].writes) clean = clean && root_of(P, t) != a_root;
      st.early_a = clean;
    }
    prev = (int)s;
  }
  return YOTA_OK;
}

static int plan_memory(yota_program& P) {
  const int ns = (int)P.steps.size();
  // constants first
  size_t coff = 0;
  for (size_t t = 0; t < P.T.size(); t++)
    if (P.T[t].d.kind == YOTA_T_CONST || P.T[t].d.kind == YOTA_T_SEED) {
      P.T[t].const_index = (int)P.const_tensors.size();
      P.const_tensors.push_back((int)t);
      coff += 256;  // one 256-byte line per scalar keeps every operand 16-byte aligned
    }
  P.counter_off = coff;
  coff += (P.counter_bytes + 255) & ~(size_t)255;
  P.chain_flags_off = coff;
  coff += (size_t)P.n_chains * 256;
  P.const_bytes = coff;
  // last use per root tensor
  for (int s = 0; s < ns; s++) {
    const Step& st = P.steps[s];
    const int until = st.fork >= 0 ? ns - 1 : s;  // a forked leaf may still be reading when the sweep ends
    for (int t : step_reads(P, st)) {
      TInfo& r = P.T[root_of(P, t)];
      r.last_step = std::max(r.last_step, until);
    }
  }
  for (size_t t = 0; t < P.T.size(); t++)
    if (P.T[t].alias_of >= 0 && P.T[t].d.kind == YOTA_T_OUTPUT && P.T[t].needed) {
      TInfo& r = P.T[root_of(P, (int)t)];
      if (r.d.kind == YOTA_T_TEMP) r.fwd_out_slot = P.T[t].d.slot;
    }
  std::vector<std::vector<int>> defs(ns), frees(ns);
  for (size_t t = 0; t < P.T.size(); t++) {
    TInfo& ti = P.T[t];
    if (ti.alias_of >= 0 || ti.d.kind != YOTA_T_TEMP || !ti.materialized || ti.def_step < 0) continue;
    if (ti.fwd_out_slot >= 0) continue;  // lives in the caller's output buffer
    if (ti.last_step < ti.def_step) ti.last_step = ti.def_step;
    defs[ti.alloc_step >= 0 ? ti.alloc_step : ti.def_step].push_back((int)t);
    frees[ti.last_step].push_back((int)t);
  }
  FreeList fl;
  fl.end = P.const_bytes;
  // workspaces: allocated at their owner step, released after the finisher / same step
  std::vector<std::vector<std::pair<size_t, size_t>>> ws_frees(ns);
  for (int s = 0; s < ns; s++) {  // sorted on the side stream from the start of the run: live from step 0
    Step& st = P.steps[s];
    if (st.kind == STEP_UNGETINDEX && st.ws_bytes > 0 && st.presort >= 0) {
      st.ws_off = (long long)fl.alloc(st.ws_bytes);
      ws_frees[s].push_back({(size_t)st.ws_off, st.ws_bytes});
    }
  }
  for (int s = 0; s < ns; s++) {
    Step& st = P.steps[s];
    for (auto& sh : P.shadows)
      if (sh.first_step == s) sh.arena_off = (long long)fl.alloc((size_t)P.T[sh.tensor].numel * P.shadow_elem_bytes);
    for (auto& sh : P.tshadows)
      if (sh.first_step == s) sh.arena_off = (long long)fl.alloc((size_t)P.T[sh.tensor].numel * 4);
    for (int t : defs[s]) P.T[t].arena_off = (long long)fl.alloc((size_t)P.T[t].numel * elem_size(P.T[t]));
    if (st.kind == STEP_REGION || (st.kind == STEP_GEMM && st.epi_region >= 0)) {
      Region& r = P.regions[st.kind == STEP_REGION ? st.region : st.epi_region];
      const int n_store = r.L.p.n_out - (int)r.sums.size();
      for (size_t q = 0; q < r.sums.size(); q++) {
        Step& fs = P.steps[r.finish_steps[q]];
        fs.ws_off = (long long)fl.alloc(fs.ws_bytes);
        r.ws_off[n_store + q] = fs.ws_off;
        ws_frees[fs.fork >= 0 ? ns - 1 : r.finish_steps[q]].push_back({(size_t)fs.ws_off, fs.ws_bytes});
      }
    } else if ((st.kind == STEP_UNGETINDEX || st.kind == STEP_GEMM) && st.ws_bytes > 0 && st.presort < 0) {
      st.ws_off = (long long)fl.alloc(st.ws_bytes);
      ws_frees[s].push_back({(size_t)st.ws_off, st.ws_bytes});
    }
    for (int t : frees[s]) fl.release((size_t)P.T[t].arena_off, (size_t)P.T[t].numel * elem_size(P.T[t]));
    for (auto& w : ws_frees[s]) fl.release(w.first, w.second);
    for (auto& sh : P.shadows)
      if (sh.last_step == s) fl.release((size_t)sh.arena_off, (size_t)P.T[sh.tensor].numel * P.shadow_elem_bytes);
    for (auto& sh : P.tshadows)
      if (sh.last_step == s) fl.release((size_t)sh.arena_off, (size_t)P.T[sh.tensor].numel * 4);
  }
  P.arena_bytes = fl.end;
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// index spec from op attributes
// ------------------------------------------------------------------------------------------
static int index_spec_of(const yota_program& P, const yota_op_desc& o, const TInfo& x, IndexSpec& sp,
                         long long* n_sel) {
  memset(&sp, 0, sizeof(sp));
  sp.n = (int)o.iattr[0];
  if (sp.n < 1 || sp.n > 4) YOTA_FAIL(YOTA_E_SHAPE, "getindex: 1..4 indices supported");
  if (sp.n == 1) {
    sp.xdims[0] = x.numel;  // linear indexing
  } else {
    if (sp.n < x.d.ndim) YOTA_FAIL(YOTA_E_SHAPE, "getindex: %d indices for a %d-d array", sp.n, x.d.ndim);
    for (int d = 0; d < sp.n; d++) sp.xdims[d] = d < x.d.ndim ? x.d.dims[d] : 1;
  }
  long long nsel = 1;
  for (int d = 0; d < sp.n; d++) {
    sp.kind[d] = (int)o.iattr[1 + 3 * d];
    sp.a[d] = o.iattr[2 + 3 * d];
    sp.b[d] = o.iattr[3 + 3 * d];
    long long ext = 1;
    switch (sp.kind[d]) {
      case YOTA_IDX_COLON:
        sp.a[d] = 1;
        sp.b[d] = sp.xdims[d];
        ext = sp.xdims[d];
        break;
      case YOTA_IDX_INT:
        sp.b[d] = sp.a[d];
        if (sp.a[d] < 1 || sp.a[d] > sp.xdims[d]) YOTA_FAIL(YOTA_E_SHAPE, "getindex: BoundsError");
        break;
      case YOTA_IDX_RANGE:
        if (sp.a[d] < 1 || sp.b[d] > sp.xdims[d]) YOTA_FAIL(YOTA_E_SHAPE, "getindex: BoundsError");
        ext = std::max(0ll, sp.b[d] - sp.a[d] + 1);
        break;
      case YOTA_IDX_VEC: {
        const int slot = (int)sp.a[d];
        if (1 + slot >= o.n_in) YOTA_FAIL(YOTA_E_ARG, "getindex: missing index-vector operand");
        const TInfo& v = P.T[o.in[1 + slot]];
        if (v.d.dtype != YOTA_I64) YOTA_FAIL(YOTA_E_ARG, "getindex: index vectors must be Int64");
        sp.vlen[d] = v.numel;
        ext = v.numel;
        break;
      }
      default:
        YOTA_FAIL(YOTA_E_ARG, "getindex: bad index kind");
    }
    nsel *= ext;
  }
  sp.nd_x = sp.n;
  *n_sel = nsel;
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// unrolled loops: batching of the time-independent contractions
// ------------------------------------------------------------------------------------------
// Umlaut unrolls `for t in 1:T` at trace time, so an Elman RNN arrives as
// T copies of  Wx*x[:,:,t],  and its pullback as T-term accumulation chains
//     dWx = sum_t dz_t * x[:,:,t]',  dWh = sum_t dz_t * h_{t-1}',  db = sum_t sum(dz_t; dims=2),
//     dx  = sum_t ungetindex(Wx' * dz_t, :, :, t)
// built from `broadcast(+, dx, old)` ops (src/grad.jl:83-95).  Only  Wh*h_{t-1}  and  Wh'*dz_t
// are truly serial.  This pass rewrites the op list (before DCE / fusion):
//   * the T forward products become ONE  Wx * reshape(x, D, B*T)  whose result is read through
//     slice views;
//   * every accumulation chain whose leaves are per-step contractions of members of one family
//     becomes ONE contraction over the concatenation  [dz_1 ... dz_T]  (K = T*B) -- and the
//     concatenation costs nothing: the members are PLACED in one buffer (internal op OP_CATV:
//     each member becomes a view of the container, written there by its own producer).
// The summation order of the batched sums differs from the tape's left fold: this is the rtol
// contract of the contraction modes (never applied to index-vector scatter-adds, whose pullbacks
// are bit-exact), SURVEY §7 "Tape accumulation" (iii).  C4: 1290 -> ~520 steps.
constexpr int OP_CATV = 200;

static void init_tinfo(TInfo& ti) {
  ti.numel = 1;
  int nd = ti.d.ndim;
  for (int d = 0; d < ti.d.ndim; d++) ti.numel *= ti.d.dims[d];
  while (nd > 0 && ti.d.dims[nd - 1] == 1) nd--;
  ti.nd = nd;
  for (int d = 0; d < 4; d++) ti.dims[d] = d < nd ? ti.d.dims[d] : 1;
}

static void link_ops(yota_program& P) {
  for (auto& t : P.T) {
    t.producer = -1;
    t.consumers.clear();
  }
  for (size_t i = 0; i < P.ops.size(); i++) {
    const yota_op_desc& o = P.ops[i];
    P.T[o.out].producer = (int)i;
    for (int j = 0; j < o.n_in; j++) P.T[o.in[j]].consumers.push_back((int)i);
    if (o.opcode == OP_CATV)
      for (int m : P.catv[o.out]) P.T[m].consumers.push_back((int)i);
  }
}

static int add_tensor(yota_program& P, std::initializer_list<long long> dims) {
  TInfo t;
  memset(&t.d, 0, sizeof(t.d));
  t.d.dtype = YOTA_F32;
  t.d.kind = YOTA_T_TEMP;
  t.d.slot = -1;
  t.d.ndim = (int)dims.size();
  int k = 0;
  for (long long v : dims) t.d.dims[k++] = v;
  init_tinfo(t);
  P.T.push_back(t);
  return (int)P.T.size() - 1;
}

static yota_op_desc mk_op(int opcode, std::initializer_list<int> ins, int out) {
  yota_op_desc o;
  memset(&o, 0, sizeof(o));
  o.opcode = opcode;
  o.n_in = (int)ins.size();
  int k = 0;
  for (int v : ins) o.in[k++] = v;
  o.out = out;
  return o;
}

static bool same_shape(const TInfo& a, const TInfo& b) {
  if (a.nd != b.nd) return false;
  for (int d = 0; d < a.nd; d++)
    if (a.dims[d] != b.dims[d]) return false;
  return true;
}

static void rewrite_unrolled_loops(yota_program& P) {
  const int mode = gemm_mode_of(P.flags);
  // (BF16 / 3xTF32 keep whole-array operand shadows that are cast at first use: a container whose
  //  members appear over time cannot be shadowed that way)
  if ((P.flags & YOTA_FLAG_NO_FUSE) || mode == 2 || mode == 3 || getenv("YOTA_NO_LOOP_BATCHING")) return;
  const int n_ops = (int)P.ops.size();
  auto uses = [&](int t) { return (int)P.T[t].consumers.size(); };
  struct Sl {
    int X;
    long long t;
  };
  auto slice_of = [&](int t, Sl* s) {
    const int p = P.T[t].producer;
    if (p < 0) return false;
    const yota_op_desc& o = P.ops[p];
    if (o.opcode != YOTA_OP_GETINDEX || o.iattr[0] != 3) return false;
    const TInfo& x = P.T[o.in[0]];
    if (x.d.ndim != 3 || x.d.dtype != YOTA_F32) return false;
    if (o.iattr[1] != YOTA_IDX_COLON || o.iattr[4] != YOTA_IDX_COLON || o.iattr[7] != YOTA_IDX_INT) return false;
    s->X = o.in[0];
    s->t = o.iattr[8];
    return s->t >= 1 && s->t <= x.d.dims[2];
  };
  std::map<int, std::vector<yota_op_desc>> ins_before, ins_after;
  std::map<int, yota_op_desc> replace;
  std::map<int, int> x2_of;  // 3-d array -> its (D, B*T) reshape (made by the forward batching, reused by dWx)
  int n_batched = 0;

  // ---- forward: A * x[:, :, t] for all t of an INPUT array -> one A * reshape(x) ----------------
  {
    std::map<std::pair<int, int>, std::vector<std::pair<long long, int>>> fam;
    for (int i = 0; i < n_ops; i++) {
      const yota_op_desc& o = P.ops[i];
      Sl s;
      if (o.opcode != YOTA_OP_MATMUL || o.iattr[0] || o.iattr[1] || !slice_of(o.in[1], &s)) continue;
      if (P.T[o.in[0]].d.kind != YOTA_T_INPUT || P.T[s.X].d.kind != YOTA_T_INPUT || P.T[o.in[0]].d.ndim != 2) continue;
      fam[{o.in[0], s.X}].push_back({s.t, i});
    }
    for (auto& kv : fam) {
      auto& v = kv.second;
      const TInfo &A = P.T[kv.first.first], &X = P.T[kv.first.second];
      const long long D = X.d.dims[0], B = X.d.dims[1], T = X.d.dims[2], M = A.d.dims[0];
      if ((long long)v.size() != T || T < 2 || A.d.dims[1] != D || (M * B) % 4) continue;
      std::sort(v.begin(), v.end());
      bool ok = true;
      int first = n_ops;
      for (long long k = 0; k < T; k++) {
        ok = ok && v[k].first == k + 1;
        first = std::min(first, v[k].second);
      }
      if (!ok) continue;
      const int X2 = add_tensor(P, {D, B * T}), Zall = add_tensor(P, {M, B * T}), Z3 = add_tensor(P, {M, B, T});
      x2_of[kv.first.second] = X2;
      auto& pre = ins_before[first];
      pre.push_back(mk_op(YOTA_OP_RESHAPE, {kv.first.second}, X2));
      pre.push_back(mk_op(YOTA_OP_MATMUL, {kv.first.first, X2}, Zall));
      pre.push_back(mk_op(YOTA_OP_RESHAPE, {Zall}, Z3));
      for (auto& m : v) {
        yota_op_desc g = mk_op(YOTA_OP_GETINDEX, {Z3}, P.ops[m.second].out);
        g.iattr[0] = 3;
        g.iattr[1] = YOTA_IDX_COLON;
        g.iattr[4] = YOTA_IDX_COLON;
        g.iattr[7] = YOTA_IDX_INT;
        g.iattr[8] = g.iattr[9] = m.first;
        replace[m.second] = g;
      }
      n_batched++;
    }
  }

  // ---- pullback: accumulation chains  acc = leaf_n + (... + (leaf_1 + leaf_0)) ------------------
  struct Chain {
    std::vector<int> leaves, adds;
    int tail;
  };
  std::vector<Chain> chains;
  {
    std::vector<int> tail_of(P.T.size(), -1);
    for (int i = 0; i < n_ops; i++) {
      const yota_op_desc& o = P.ops[i];
      if (o.opcode != YOTA_EW_ADD || o.in[0] == o.in[1]) continue;
      const int a = o.in[0], b = o.in[1];
      if (!same_shape(P.T[a], P.T[o.out]) || !same_shape(P.T[b], P.T[o.out])) continue;
      int c = -1, leaf = -1;
      if (tail_of[a] >= 0 && uses(a) == 1) c = tail_of[a], leaf = b;
      else if (tail_of[b] >= 0 && uses(b) == 1) c = tail_of[b], leaf = a;
      if (c >= 0) {
        chains[c].leaves.push_back(leaf);
        chains[c].adds.push_back(i);
        chains[c].tail = o.out;
      } else {
        chains.push_back({{a, b}, {i}, o.out});
        c = (int)chains.size() - 1;
      }
      tail_of[o.out] = c;
    }
  }
  enum { NT_SLICE, NT_PLAIN, UNGET_TN, SUMK };
  struct Leaf {
    int kind, dz, other;  // other: X (NT_SLICE), h (NT_PLAIN), W (UNGET_TN)
    long long t;
  };
  auto classify = [&](int l, Leaf* lf) {
    if (uses(l) != 1 || P.T[l].producer < 0) return false;
    const yota_op_desc& o = P.ops[P.T[l].producer];
    Sl s;
    if (o.opcode == YOTA_OP_MATMUL && !o.iattr[0] && o.iattr[1]) {
      if (slice_of(o.in[1], &s))
        *lf = {NT_SLICE, o.in[0], s.X, s.t};
      else
        *lf = {NT_PLAIN, o.in[0], o.in[1], 0};
      return true;
    }
    if (o.opcode == YOTA_OP_UNGETINDEX && o.iattr[0] == 3 && o.iattr[1] == YOTA_IDX_COLON && o.iattr[4] == YOTA_IDX_COLON &&
        o.iattr[7] == YOTA_IDX_INT && o.n_in == 1) {
      const int src = o.in[0];
      if (uses(src) != 1 || P.T[src].producer < 0) return false;
      const yota_op_desc& m = P.ops[P.T[src].producer];
      if (m.opcode != YOTA_OP_MATMUL || !m.iattr[0] || m.iattr[1]) return false;
      *lf = {UNGET_TN, m.in[1], m.in[0], o.iattr[8]};
      return true;
    }
    if (o.opcode == YOTA_OP_SUM && o.iattr[0] == 2 && o.fattr[0] == 1.0 && P.T[o.in[0]].d.ndim == 2) {
      *lf = {SUMK, o.in[0], -1, 0};
      return true;
    }
    return false;
  };
  // containers (zero-copy concatenations) made so far
  struct Cont {
    int tensor;
    std::vector<int> members;
  };
  std::vector<Cont> conts;
  std::map<int, std::pair<int, int>> where;  // member -> (container, position)
  auto placeable = [&](int m, const TInfo& like) {
    const TInfo& t = P.T[m];
    if (t.d.kind != YOTA_T_TEMP || t.d.dtype != YOTA_F32 || t.d.ndim != 2 || t.producer < 0) return false;
    if (t.d.dims[0] != like.d.dims[0] || t.d.dims[1] != like.d.dims[1] || (t.numel % 4)) return false;
    const int opc = P.ops[t.producer].opcode;
    return is_ew(opc) || opc == YOTA_OP_MATMUL;  // written whole by its own kernel (not a view)
  };
  // 0: a new container can be made, 1: a view of an existing one, -1: neither
  auto contain_kind = [&](const std::vector<int>& ms) {
    int in = 0;
    for (int m : ms) in += where.count(m) ? 1 : 0;
    if (in == 0) {
      std::vector<int> sorted = ms;
      std::sort(sorted.begin(), sorted.end());
      if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) return -1;
      for (int m : ms)
        if (!placeable(m, P.T[ms[0]])) return -1;
      return 0;
    }
    if (in != (int)ms.size()) return -1;
    const auto w0 = where[ms[0]];
    for (size_t k = 0; k < ms.size(); k++)
      if (where[ms[k]] != std::make_pair(w0.first, w0.second + (int)k)) return -1;
    return 1;
  };
  // tensor holding [ms...] side by side; new ops go before op `at` (views) / after the last producer (containers)
  auto container_of = [&](const std::vector<int>& ms, int at) {
    const TInfo& m0 = P.T[ms[0]];
    const long long M = m0.d.dims[0], B = m0.d.dims[1], n = (long long)ms.size();
    if (contain_kind(ms) == 0) {
      const int ct = add_tensor(P, {M, B * n});
      P.catv[ct] = ms;
      int last = 0;
      for (size_t k = 0; k < ms.size(); k++) {
        where[ms[k]] = {(int)conts.size(), (int)k};
        last = std::max(last, P.T[ms[k]].producer);
      }
      conts.push_back({ct, ms});
      ins_after[last].push_back(mk_op(OP_CATV, {}, ct));
      return ct;
    }
    const auto w0 = where[ms[0]];
    const Cont& c = conts[w0.first];
    if (w0.second == 0 && ms.size() == c.members.size()) return c.tensor;
    // a column range of the container: usable as soon as ITS members exist (the chain that sums
    // dz_2 .. dz_T closes before dz_1 -- and with it the whole container -- has been produced)
    (void)at;
    const int v = add_tensor(P, {M, B * n});
    P.catv[v] = ms;
    P.catv_view[v] = {c.tensor, (long long)w0.second * M * B};
    int last = 0;
    for (int m : ms) last = std::max(last, P.T[m].producer);
    ins_after[last].push_back(mk_op(OP_CATV, {}, v));
    return v;
  };
  for (int round = 0; round < 2; round++)
    for (Chain& ch : chains) {
      if (ch.leaves.size() < 3 || ch.tail < 0) continue;
      std::vector<Leaf> lf(ch.leaves.size());
      bool ok = true;
      for (size_t k = 0; k < ch.leaves.size() && ok; k++) ok = classify(ch.leaves[k], &lf[k]) && lf[k].kind == lf[0].kind;
      if (!ok) continue;
      const int kind = lf[0].kind, at = ch.adds.back();
      if ((round == 0) != (kind == NT_SLICE || kind == UNGET_TN)) continue;
      for (auto& l : lf)
        ok = ok && (kind == NT_PLAIN || l.other == lf[0].other) && P.T[l.dz].d.ndim == 2 && same_shape(P.T[l.dz], P.T[lf[0].dz]);
      if (!ok) continue;
      std::vector<int> dzs;
      yota_op_desc rep;
      if (kind == NT_SLICE || kind == UNGET_TN) {
        std::sort(lf.begin(), lf.end(), [](const Leaf& a, const Leaf& b) { return a.t < b.t; });
        for (size_t k = 0; k < lf.size(); k++) ok = ok && lf[k].t == lf[0].t + (long long)k;
        for (auto& l : lf) dzs.push_back(l.dz);
        if (!ok || contain_kind(dzs) < 0) continue;
        if (kind == NT_SLICE) {
          const TInfo& X = P.T[lf[0].other];
          const long long D = X.d.dims[0], B = X.d.dims[1], T = X.d.dims[2], n = (long long)lf.size();
          if (P.T[dzs[0]].d.dims[1] != B) continue;
          const int dZ = container_of(dzs, at);
          int X2;
          if (x2_of.count(lf[0].other) && X.d.kind == YOTA_T_INPUT)
            X2 = x2_of[lf[0].other];
          else {
            X2 = add_tensor(P, {D, B * T});
            ins_before[at].push_back(mk_op(YOTA_OP_RESHAPE, {lf[0].other}, X2));
          }
          int Bop = X2;
          if (n != T) {
            Bop = add_tensor(P, {D, B * n});
            yota_op_desc g = mk_op(YOTA_OP_GETINDEX, {X2}, Bop);
            g.iattr[0] = 2;
            g.iattr[1] = YOTA_IDX_COLON;
            g.iattr[4] = YOTA_IDX_RANGE;
            g.iattr[5] = (lf[0].t - 1) * B + 1;
            g.iattr[6] = (lf[0].t - 1 + n) * B;
            ins_before[at].push_back(g);
          }
          rep = mk_op(YOTA_OP_MATMUL, {dZ, Bop}, ch.tail);
          rep.iattr[1] = 1;
        } else {
          const TInfo& out = P.T[ch.tail];
          const long long T = out.d.ndim == 3 ? out.d.dims[2] : 0;
          if ((long long)lf.size() != T || lf[0].t != 1 || P.T[dzs[0]].d.dims[1] != out.d.dims[1]) continue;
          const int dZ = container_of(dzs, at);
          const int mm = add_tensor(P, {out.d.dims[0], out.d.dims[1] * T});
          yota_op_desc g = mk_op(YOTA_OP_MATMUL, {lf[0].other, dZ}, mm);
          g.iattr[0] = 1;
          ins_before[at].push_back(g);
          rep = mk_op(YOTA_OP_RESHAPE, {mm}, ch.tail);
        }
      } else {
        // the members must already sit side by side in a container (made by a chain that knows t)
        for (auto& l : lf) ok = ok && where.count(l.dz);
        if (!ok) continue;
        std::sort(lf.begin(), lf.end(), [&](const Leaf& a, const Leaf& b) { return where[a.dz] < where[b.dz]; });
        for (auto& l : lf) dzs.push_back(l.dz);
        if (contain_kind(dzs) != 1) continue;
        if (kind == NT_PLAIN) {
          std::vector<int> hs;
          for (auto& l : lf) hs.push_back(l.other);
          ok = true;  // every leaf has its own h: same shapes, one column block per dz
          for (int h : hs) ok = ok && P.T[h].d.ndim == 2 && same_shape(P.T[h], P.T[hs[0]]) && P.T[h].d.dims[1] == P.T[dzs[0]].d.dims[1];
          if (!ok || contain_kind(hs) < 0) continue;
          const int dZ = container_of(dzs, at), H = container_of(hs, at);
          rep = mk_op(YOTA_OP_MATMUL, {dZ, H}, ch.tail);
          rep.iattr[1] = 1;
        } else {
          const int dZ = container_of(dzs, at);
          rep = mk_op(YOTA_OP_SUM, {dZ}, ch.tail);
          rep.iattr[0] = 2;
          rep.fattr[0] = 1.0;
        }
      }
      replace[at] = rep;
      ch.tail = -1;  // done
      n_batched++;
    }
  if (!n_batched) return;
  std::vector<yota_op_desc> ops;
  for (int i = 0; i < n_ops; i++) {
    for (auto& o : ins_before[i]) ops.push_back(o);
    ops.push_back(replace.count(i) ? replace[i] : P.ops[i]);
    for (auto& o : ins_after[i]) ops.push_back(o);
  }
  P.ops.swap(ops);
  P.n_batched = n_batched;
  link_ops(P);
}

// Peephole on the op list (before dead-code elimination): the scatter-add pullback of
// `sum(x[:, idx] .* V)` arrives as  D = copy(seed) (the broadcast-back of the sum's pullback),
// dG = D .* V,  dx = ungetindex(dG, :, idx).  dG exists only to be scattered, so the fold reads V
// in place and multiplies by the seed on load -- the same single rounding per product, the same
// fold order, bit-identical -- and the 4 GiB product (C5) is neither written nor read back.
static void rewrite_scaled_scatter(yota_program& P) {
  P.op_scale.assign(P.ops.size(), -1);
  if ((P.flags & YOTA_FLAG_NO_FUSE) || getenv("YOTA_NO_SCALED_SCATTER")) return;
  for (size_t u = 0; u < P.ops.size(); u++) {
    yota_op_desc& U = P.ops[u];
    if (U.opcode != YOTA_OP_UNGETINDEX || U.iattr[0] != 2) continue;
    if (U.iattr[1] != YOTA_IDX_COLON || U.iattr[4] != YOTA_IDX_VEC) continue;
    const int t = U.in[0];
    const TInfo& tt = P.T[t];
    if (tt.d.kind != YOTA_T_TEMP || tt.producer < 0 || tt.consumers.size() != 1 || tt.d.ndim != 2) continue;
    const yota_op_desc& M = P.ops[tt.producer];
    if (M.opcode != YOTA_EW_MUL) continue;
    for (int side = 0; side < 2; side++) {
      const int d = M.in[side], v = M.in[1 - side];
      const TInfo &td = P.T[d], &tv = P.T[v];
      if (td.producer < 0 || P.ops[td.producer].opcode != YOTA_EW_COPY) continue;
      const int sc = P.ops[td.producer].in[0];
      const TInfo& ts = P.T[sc];
      if (ts.numel != 1 || (ts.d.kind != YOTA_T_SEED && ts.d.kind != YOTA_T_CONST && ts.d.kind != YOTA_T_INPUT)) continue;
      if (tv.nd != tt.nd || tv.numel != tt.numel || tv.d.dtype != YOTA_F32) continue;
      bool same = true;
      for (int q = 0; q < 4; q++) same &= tv.dims[q] == tt.dims[q];
      if (!same) continue;
      U.in[0] = v;
      P.op_scale[u] = sc;
      P.T[t].consumers.clear();
      P.T[v].consumers.push_back((int)u);
      P.T[sc].consumers.push_back((int)u);
      break;
    }
  }
}

// getindex(x, :, idx) -> consumer-region fusion: when every reader of the gathered columns is an
// elementwise op of ONE region in which it is a FULL operand, the region reads x[:, idx[c]]
// itself (operand class CLS_GATHER) and the gathered copy (C5: 4 GiB written, 4 GiB read back)
// never exists.
static int plan_gather_fusion(yota_program& P) {
  if ((P.flags & YOTA_FLAG_NO_FUSE) || getenv("YOTA_NO_GATHER_FUSION")) return YOTA_OK;
  for (size_t si = 0; si < P.steps.size(); si++) {
    Step& gs = P.steps[si];
    if (gs.kind != STEP_GETINDEX) continue;
    const yota_op_desc& o = P.ops[gs.op];
    if (o.iattr[0] != 2 || o.iattr[1] != YOTA_IDX_COLON || o.iattr[4] != YOTA_IDX_VEC) continue;
    const int x = o.in[0], t = o.out, iv = o.in[1 + (int)o.iattr[5]];
    const TInfo &xt = P.T[x], &tt = P.T[t];
    if (xt.d.ndim != 2 || tt.d.kind != YOTA_T_TEMP || tt.alias_of >= 0 || P.T[iv].d.dtype != YOTA_I64) continue;
    int ri = -1;
    bool ok = true;
    for (int c : tt.consumers) {
      if (!P.op_needed[c]) continue;
      const yota_op_desc& co = P.ops[c];
      if (!is_ew(co.opcode) || P.T[co.out].region < 0) ok = false;
      else if (ri < 0) ri = P.T[co.out].region;
      else if (ri != P.T[co.out].region) ok = false;
    }
    for (size_t q = 0; q < P.T.size(); q++)
      if (P.T[q].alias_of == t && P.T[q].needed) ok = false;
    if (!ok || ri < 0) continue;
    Region& r = P.regions[ri];
    RegionParams& rp = r.L.p;
    if (r.step <= (int)si || P.steps[r.step].kind != STEP_REGION) continue;
    if (r.k != 1 || rp.R != xt.d.dims[0] || rp.C != P.T[iv].numel) continue;
    int q_in = -1;
    for (int q = 0; q < rp.n_in; q++)
      if (r.in_tensors[q] == t && rp.in[q].cls == CLS_FULL) q_in = q;
    if (q_in < 0) continue;
    rp.in[q_in].cls = CLS_GATHER;
    r.gather_x.assign(rp.n_in, -1);
    r.gather_idx.assign(rp.n_in, -1);
    r.gather_x[q_in] = x;
    r.gather_idx[q_in] = iv;
    Step& rs = P.steps[r.step];
    for (int& rd : rs.reads)
      if (rd == t) rd = x;
    rs.reads.push_back(iv);
    P.T[root_of(P, x)].materialized = true;
    P.T[t].materialized = false;
    gs.kind = STEP_NOP;
    gs.label = "(gathered inside step " + std::to_string(r.step) + ")";
    // the operand class is part of the micro-program's signature
    if (r.L.static_id >= 0) P.n_static--;
    r.sig = region_signature(rp, r.L.V);
    r.L.static_id = (P.flags & YOTA_FLAG_NO_STATIC) ? -1 : static_chain_lookup(rp, r.L.V);
    if (r.L.static_id >= 0) P.n_static++;
    const size_t cut = rs.label.find(" static=") != std::string::npos ? rs.label.find(" static=") : rs.label.find(" interp");
    rs.label = rs.label.substr(0, cut) + " gather" + (r.L.static_id >= 0 ? std::string(" static=") + static_chain_name(r.L.static_id) : " interp");
  }
  return YOTA_OK;
}

// GEMM -> consumer-region fusion: when the only reader of a tensor-core GEMM's output is one
// fused region whose micro-program exists as an epilogue variant (bias [+ tanh]; tanh' * delta
// + db), the region runs inside the GEMM's epilogue and the GEMM output never touches HBM.
static int plan_gemm_epilogues(yota_program& P) {
  if (P.flags & YOTA_FLAG_NO_FUSE) return YOTA_OK;
  const int mode = gemm_mode_of(P.flags);
  if (mode == 0 || getenv("YOTA_NO_EPILOGUE")) return YOTA_OK;
  for (size_t si = 0; si < P.steps.size(); si++) {
    Step& gs = P.steps[si];
    if (gs.kind != STEP_GEMM || gs.acc_tensor >= 0) continue;
    const yota_op_desc& o = P.ops[gs.op];
    const int z = o.out;
    const TInfo& zt = P.T[z];
    if (zt.d.kind != YOTA_T_TEMP || zt.alias_of >= 0) continue;
    GemmArgs g{};
    g.mode = mode;
    g.transA = (int)o.iattr[0];
    g.transB = (int)o.iattr[1];
    gemm_dims(P, o, &g.M, &g.N, &g.K, &g.lda, &g.ldb);
    const bool two_cta = gemm_tc2_eligible(g, P.ctx->sm_count);
    if (!two_cta && !(mode == 1 && gemm_tc_eligible(g))) continue;
    // every needed consumer of z is an elementwise op of one region
    int ri = -1;
    bool ok = true;
    for (int c : zt.consumers) {
      if (!P.op_needed[c]) continue;
      const yota_op_desc& co = P.ops[c];
      if (!is_ew(co.opcode) || P.T[co.out].region < 0) ok = false;
      else if (ri < 0) ri = P.T[co.out].region;
      else if (ri != P.T[co.out].region) ok = false;
    }
    if (!ok || ri < 0) continue;
    for (size_t q = 0; q < P.T.size(); q++)  // an alias (reshape) of z read elsewhere keeps it in memory
      if (P.T[q].alias_of == z && P.T[q].needed) ok = false;
    Region& r = P.regions[ri];
    const RegionParams& rp = r.L.p;
    if (!ok || r.L.static_id < 0 || r.step <= (int)si || P.steps[r.step].kind != STEP_REGION) continue;
    if (r.k != 1 || rp.R != g.M || rp.C != g.N) continue;
    int z_in = -1;
    for (int q = 0; q < rp.n_in; q++) {
      if (r.in_tensors[q] == z) {
        z_in = q;
        if (rp.in[q].cls != CLS_FULL) ok = false;
      } else if (step_of_tensor(P, r.in_tensors[q]) >= (int)si) {
        ok = false;  // another operand of the region is not ready when the GEMM runs
      }
    }
    if (!ok || z_in < 0) continue;
    if (two_cta ? !gemm_epilogue_chain_ok(r.L.static_id, g.transA, g.transB, z_in)
                : !(r.sums.empty() && !getenv("YOTA_NO_EPILOGUE1") &&
                    gemm_epilogue1_chain_ok(g, P.ctx->sm_count, r.L.static_id, z_in)))
      continue;
    // absorb
    gs.epi_region = ri;
    gs.epi_z_in = z_in;
    gs.label += " + epilogue{" + P.steps[r.step].label.substr(0, P.steps[r.step].label.find(' ')) + "}";
    for (int t : r.in_tensors)
      if (t != z) gs.reads.push_back(t);
    Step& rs = P.steps[r.step];
    rs.kind = STEP_NOP;
    rs.label = "(absorbed into step " + std::to_string(si) + ")";
    rs.reads.clear();
    const int n_store = rp.n_out - (int)r.sums.size();
    for (int q = 0; q < n_store; q++) {
      P.T[r.out_tensors[q]].def_step = (int)si;
      gs.writes.push_back(r.out_tensors[q]);
    }
    rs.writes.clear();
    // (compensated mode: z stays in the arena for the length of this step -- the partial sums of
    // the K-chunks are accumulated there before the chain runs on the complete sum)
    P.T[z].materialized = mode == 3;
    // row-sum partials: one row per 256-column tile of the GEMM
    const int P_tiles = (int)(g.N / 256);
    for (size_t q = 0; q < r.sums.size(); q++) {
      Step& fs = P.steps[r.finish_steps[q]];
      if (fs.kind != STEP_ROWSUM_FINISH) return YOTA_E_ARG;
      fs.P = P_tiles;
      fs.ws_bytes = (size_t)P_tiles * g.M * sizeof(float);
      fs.ws_owner_step = (int)si;
      // the last CTA to deliver a partial also reduces them (gemm_tc.cu: Epilogue::end_tile): the
      // finisher launch leaves the critical path (8 x 10-33 us of a C3 step on 8 GPUs)
      if (q == 0 && gs.rs_finish < 0 && !getenv("YOTA_NO_FUSED_FINISH")) {
        gs.rs_finish = r.finish_steps[q];
        gs.label += " +finish";
        fs.fused_away = true;
        fs.label = "(rowsum finished inside step " + std::to_string(si) + ")";
        gs.writes.push_back(fs.out_tensor);
        P.T[fs.out_tensor].def_step = (int)si;
        P.counter_bytes = std::max(P.counter_bytes, (size_t)(g.M / 128) * sizeof(unsigned));
      }
    }
    r.step = (int)si;  // workspaces are allocated when the GEMM runs
  }
  return YOTA_OK;
}

// The index vectors of this op exist before the first step (program inputs or constants), so the
// sort of a scatter-add over them can start with the run.  YOTA_NO_PRESORT=1: keep it in the step.
static bool index_vectors_early(const yota_program& P, const yota_op_desc& o, const IndexSpec& sp) {
  static const bool off = getenv("YOTA_NO_PRESORT") != nullptr;
  if (off) return false;
  for (int d = 0; d < sp.n; d++)
    if (sp.kind[d] == YOTA_IDX_VEC) {
      const int k = P.T[o.in[1 + (int)sp.a[d]]].d.kind;
      if (k != YOTA_T_INPUT && k != YOTA_T_CONST) return false;
    }
  return true;
}

static int finish_steps(yota_program& P) {
  for (size_t s = 0; s < P.steps.size(); s++) {
    Step& st = P.steps[s];
    if (st.op < 0 || st.kind == STEP_ROWSUM_FINISH || st.kind == STEP_SCALAR_FINISH) continue;
    const yota_op_desc& o = P.ops[st.op];
    st.writes.push_back(st.out_override >= 0 ? st.out_override : o.out);
    if (st.kind == STEP_UNGETINDEX && P.op_scale[st.op] >= 0) st.reads.push_back(P.op_scale[st.op]);
    if (st.kind == STEP_GETINDEX || st.kind == STEP_UNGETINDEX || st.kind == STEP_SLICE_ACCUM) {
      IndexSpec sp;
      long long nsel;
      const TInfo& x = st.kind == STEP_GETINDEX ? P.T[o.in[0]] : P.T[o.out];
      YOTA_TRY(index_spec_of(P, o, x, sp, &nsel));
      const TInfo& sel = st.kind == STEP_GETINDEX ? P.T[o.out] : P.T[o.in[0]];
      if (sel.numel != nsel && !(st.kind != STEP_GETINDEX && sel.numel == 1))
        YOTA_FAIL(YOTA_E_SHAPE, "%s: selection has %lld elements, tensor has %lld", opname(o.opcode), nsel, sel.numel);
      if (st.kind == STEP_UNGETINDEX) {
        st.ws_bytes = ungetindex_workspace_bytes(sp);
        st.unget_launches = ungetindex_launch_count(sp);
        if (st.ws_bytes > 0 && index_vectors_early(P, o, sp)) st.presort = P.n_presort++;
      }
    }
    char buf[128];
    const TInfo& ot = P.T[o.out];
    snprintf(buf, sizeof(buf), " -> %%%d[%lld,%lld,%lld,%lld]", o.out, ot.dims[0], ot.dims[1], ot.dims[2], ot.dims[3]);
    st.label += buf;
    if (st.presort >= 0) st.label += " (index sort on the side stream from the start of the run)";
  }
  return YOTA_OK;
}

}  // namespace yota

// ==========================================================================================
// C ABI: build / describe / stats / destroy
// ==========================================================================================
extern "C" int yota_program_build(yota_ctx* ctx, const yota_op_desc* ops, int64_t n_ops,
                                  const yota_tensor_desc* tensors, int64_t n_tensors, uint32_t flags,
                                  yota_program** out) {
  if (!ctx || !out || n_ops < 0 || n_tensors <= 0 || (!ops && n_ops > 0) || !tensors)
    YOTA_FAIL(YOTA_E_ARG, "yota_program_build: bad arguments");
  yota_program* Pp = new yota_program();
  yota_program& P = *Pp;
  P.ctx = ctx;
  P.flags = flags;
  P.ops.assign(ops, ops + n_ops);
  P.T.resize((size_t)n_tensors);
  auto fail = [&](int code) {
    delete Pp;
    return code;
  };
  for (int64_t t = 0; t < n_tensors; t++) {
    TInfo& ti = P.T[(size_t)t];
    ti.d = tensors[t];
    if (ti.d.ndim < 0 || ti.d.ndim > YOTA_MAX_DIMS) {
      set_error("tensor %lld: ndim out of range", (long long)t);
      return fail(YOTA_E_SHAPE);
    }
    ti.numel = 1;
    int nd = ti.d.ndim;
    for (int d = 0; d < ti.d.ndim; d++) {
      if (ti.d.dims[d] < 0) {
        set_error("tensor %lld: negative dim", (long long)t);
        return fail(YOTA_E_SHAPE);
      }
      ti.numel *= ti.d.dims[d];
    }
    while (nd > 0 && ti.d.dims[nd - 1] == 1) nd--;
    ti.nd = nd;
    for (int d = 0; d < 4; d++) ti.dims[d] = d < nd ? ti.d.dims[d] : 1;
    if (ti.d.kind == YOTA_T_INPUT) P.n_in = std::max(P.n_in, ti.d.slot + 1);
    if (ti.d.kind == YOTA_T_OUTPUT) P.n_out = std::max(P.n_out, ti.d.slot + 1);
    if ((ti.d.kind == YOTA_T_CONST || ti.d.kind == YOTA_T_SEED) && ti.d.ndim != 0) {
      set_error("tensor %lld: constants and the seed must be 0-dim", (long long)t);
      return fail(YOTA_E_SHAPE);
    }
  }
  for (size_t i = 0; i < P.ops.size(); i++) {
    const yota_op_desc& o = P.ops[i];
    if (o.out >= 0 && o.out < n_tensors) {
      if (P.T[o.out].producer >= 0) {
        set_error("tensor %d has two writers (ops must be SSA)", o.out);
        return fail(YOTA_E_ARG);
      }
      P.T[o.out].producer = (int)i;
    }
    for (int j = 0; j < o.n_in && j < YOTA_MAX_OP_INPUTS; j++)
      if (o.in[j] >= 0 && o.in[j] < n_tensors) P.T[o.in[j]].consumers.push_back((int)i);
  }
  int st = validate(P);
  if (st) return fail(st);
  {  // batching of unrolled loops; an op list the pass got wrong (it is re-validated) is discarded
    const std::vector<yota_op_desc> ops0 = P.ops;
    const size_t nt0 = P.T.size();
    rewrite_unrolled_loops(P);
    if (getenv("YOTA_DEBUG_REWRITE")) fprintf(stderr, "rewrite_unrolled_loops: %d batched, %zu ops\n", P.n_batched, P.ops.size());
    if (P.n_batched && validate(P) != YOTA_OK) {
      if (getenv("YOTA_DEBUG_REWRITE")) fprintf(stderr, "  rewritten program rejected: %s\n", yota_last_error());
      P.ops = ops0;
      P.T.resize(nt0);
      P.catv.clear();
      P.catv_view.clear();
      P.n_batched = 0;
      link_ops(P);
    }
  }
  rewrite_scaled_scatter(P);
  // dead-code elimination
  P.op_needed.assign(P.ops.size(), 0);
  for (int i = (int)P.ops.size() - 1; i >= 0; i--) {
    const yota_op_desc& o = P.ops[i];
    bool need = P.T[o.out].d.kind == YOTA_T_OUTPUT;
    for (int c : P.T[o.out].consumers) need |= (bool)P.op_needed[c];
    P.op_needed[i] = need;
    P.T[o.out].needed = need;
  }
  for (auto& t : P.T)
    if (t.d.kind == YOTA_T_OUTPUT && t.producer < 0) {
      set_error("an OUTPUT tensor has no producer");
      return fail(YOTA_E_ARG);
    }
  // A scatter-add whose sort runs beside the forward sweep competes with the (persistent,
  // slot-filling) elementwise regions for CTA slots.  Measured on C5: regions sized for 8 / 7 / 6
  // CTAs per SM give 2.88 / 2.96 / 3.13 ms per eval (2.99 with the sort inside the step): the
  // gather needs all its threads to cover HBM latency, so the default stays 8; the knob is for A/B.
  for (size_t i = 0; i < P.ops.size(); i++) {
    const yota_op_desc& o = P.ops[i];
    if (!P.op_needed[i] || o.opcode != YOTA_OP_UNGETINDEX) continue;
    IndexSpec sp;
    long long nsel;
    if (index_spec_of(P, o, P.T[o.out], sp, &nsel) != YOTA_OK) continue;  // reported by finish_steps
    bool vec = false;
    for (int d = 0; d < sp.n; d++) vec |= sp.kind[d] == YOTA_IDX_VEC;
    if (vec && index_vectors_early(P, o, sp)) {
      const char* e = getenv("YOTA_PRESORT_REGION_PER_SM");
      P.region_per_sm = e && atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : 8;
      break;
    }
  }
  if ((st = plan_fusion(P))) return fail(st);
  for (size_t r = 0; r < P.regions.size(); r++)
    if ((st = codegen_region(P, (int)r))) return fail(st);
  if ((st = plan_gather_fusion(P))) return fail(st);
  if ((st = plan_gemm_epilogues(P))) return fail(st);
  if ((st = finish_steps(P))) return fail(st);
  if ((st = plan_bf16_shadows(P))) return fail(st);
  if ((st = plan_transposed_shadows(P))) return fail(st);
  if ((st = plan_recurrence_chains(P))) return fail(st);
  if ((st = plan_split_k(P))) return fail(st);
  if ((st = plan_side_leaves(P))) return fail(st);
  if ((st = plan_dependent_launch(P))) return fail(st);
  if ((st = plan_memory(P))) return fail(st);
  P.n_launches = 0;
  for (auto& s : P.steps)  // absorbed regions launch nothing; BF16 operand casts are launches of their GEMM step
    P.n_launches += (s.kind == STEP_NOP || s.fused_away) ? 0 : (s.kind == STEP_UNGETINDEX ? s.unget_launches : 1) + (s.ksplit > 1 ? 1 : 0) + (s.cast[0] ? 1 : 0) + (s.cast[1] ? 1 : 0);
  *out = Pp;
  return YOTA_OK;
}

extern "C" int64_t yota_program_describe(yota_program* p, char* buf, int64_t cap) {
  if (!p) return 0;
  std::string s;
  char line[512];
  snprintf(line, sizeof(line), "program: %zu ops, %zu steps, %zu regions (%d static), arena %zu bytes\n", p->ops.size(),
           p->steps.size(), p->regions.size(), p->n_static, p->arena_bytes);
  s += line;
  if (p->n_batched) {
    snprintf(line, sizeof(line), "  unrolled loops: %d slice families / accumulation chains batched into single contractions\n", p->n_batched);
    s += line;
  }
  for (size_t i = 0; i < p->steps.size(); i++) {
    snprintf(line, sizeof(line), "  step %zu: %s\n", i, p->steps[i].label.c_str());
    s += line;
    if (p->steps[i].kind == STEP_REGION) s += "    sig " + p->regions[p->steps[i].region].sig + "\n";
  }
  if (buf && cap > 0) {
    const size_t n = std::min((size_t)cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size();
}

// The op list the planner actually works on (after the rewrite passes), as text:
//   T id dtype ndim d0 d1 d2 d3 kind slot cval
//   O opcode out n_in in... | iattr[16] | fattr[2] | members of a concatenation...
// Lets the CPU-side tests run the REWRITTEN program through their NumPy interpreter.
extern "C" int64_t yota_program_dump_ops(yota_program* p, char* buf, int64_t cap) {
  if (!p) return 0;
  std::string s;
  char line[1024];
  for (size_t t = 0; t < p->T.size(); t++) {
    const yota_tensor_desc& d = p->T[t].d;
    snprintf(line, sizeof(line), "T %zu %d %d %lld %lld %lld %lld %d %d %.17g\n", t, d.dtype, d.ndim, (long long)d.dims[0],
             (long long)d.dims[1], (long long)d.dims[2], (long long)d.dims[3], d.kind, d.slot, d.cval);
    s += line;
  }
  for (size_t i = 0; i < p->ops.size(); i++) {
    const yota_op_desc& o = p->ops[i];
    std::string l = "O " + std::to_string(o.opcode) + " " + std::to_string(o.out) + " " + std::to_string(o.n_in);
    for (int j = 0; j < o.n_in; j++) l += " " + std::to_string(o.in[j]);
    l += " |";
    for (int j = 0; j < YOTA_N_IATTR; j++) l += " " + std::to_string((long long)o.iattr[j]);
    snprintf(line, sizeof(line), " | %.17g %.17g |", o.fattr[0], o.fattr[1]);
    l += line;
    auto it = p->catv.find(o.out);
    if (o.opcode == 200 && it != p->catv.end())
      for (int m : it->second) l += " " + std::to_string(m);
    s += l + "\n";
  }
  if (buf && cap > 0) {
    const size_t n = std::min((size_t)cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size();
}

extern "C" int yota_program_stats(yota_program* p, int64_t* n_steps, int64_t* n_launches, int64_t* workspace_bytes,
                                  int64_t* n_fused_regions, int64_t* n_static_kernels) {
  if (!p) YOTA_FAIL(YOTA_E_ARG, "null program");
  if (n_steps) *n_steps = (int64_t)p->steps.size();
  if (n_launches) {
    *n_launches = p->n_launches;
    for (size_t c = 0; c < p->chain_lens.size(); c++)  // a recurrence launched persistently is one kernel
      if (p->chain_taken[c]) *n_launches -= p->chain_lens[c] - 1;
  }
  if (workspace_bytes) *workspace_bytes = (int64_t)p->arena_bytes;
  if (n_fused_regions) *n_fused_regions = (int64_t)p->regions.size();
  if (n_static_kernels) *n_static_kernels = p->n_static;
  return YOTA_OK;
}

static void free_adam_state(yota_program* p) {
  for (float* q : p->adam_m) cudaFree(q);
  for (float* q : p->adam_v) cudaFree(q);
  if (p->adam_t) cudaFree(p->adam_t);
  p->adam_m.clear();
  p->adam_v.clear();
  p->adam_t = nullptr;
  p->adam = false;
}

extern "C" int yota_program_destroy(yota_program* p) {
  if (!p) return YOTA_OK;
  for (auto& g : p->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (auto e : p->ar_events) cudaEventDestroy(e);
  for (auto e : p->fork_ev) cudaEventDestroy(e);
  for (auto e : p->join_ev) cudaEventDestroy(e);
  if (p->presort_fork) cudaEventDestroy(p->presort_fork);
  for (auto e : p->presort_join) cudaEventDestroy(e);
  free_adam_state(p);
  if (p->arena) cudaFree(p->arena);
  delete p;
  return YOTA_OK;
}

// ==========================================================================================
// executor
// ==========================================================================================
namespace yota {

static void* root_ptr(yota_program& P, int r) {
  const TInfo& ti = P.T[r];
  switch (ti.d.kind) {
    case YOTA_T_INPUT: return P.bound_in[ti.d.slot];
    case YOTA_T_OUTPUT: return P.bound_out[ti.d.slot];
    case YOTA_T_CONST:
    case YOTA_T_SEED: return P.arena + (size_t)ti.const_index * 256;
    default:
      if (ti.fwd_out_slot >= 0) return P.bound_out[ti.fwd_out_slot];
      return ti.arena_off >= 0 ? P.arena + ti.arena_off : nullptr;
  }
}
static void* tensor_ptr(yota_program& P, int t) {
  const int r = root_of(P, t);
  char* p = static_cast<char*>(root_ptr(P, r));
  const long long off = alias_offset(P, t);
  return p && off ? p + off * (P.T[r].d.dtype == YOTA_I64 ? 8 : 4) : p;
}

static int ensure_arena(yota_program& P) {
  if (P.arena) return YOTA_OK;
  YOTA_TRY(ctx_ensure_device(P.ctx));
  const size_t bytes = std::max<size_t>(P.arena_bytes, 256);
  cudaError_t e = cudaMalloc((void**)&P.arena, bytes);
  if (e != cudaSuccess) {
    P.arena = nullptr;
    YOTA_FAIL(YOTA_E_NOMEM, "cannot allocate the program arena (%zu bytes): %s", bytes, cudaGetErrorString(e));
  }
  if (P.counter_bytes) YOTA_CUDA(cudaMemsetAsync(P.arena + P.counter_off, 0, P.counter_bytes, P.ctx->stream));
  for (size_t i = 0; i < P.const_tensors.size(); i++) {
    const TInfo& t = P.T[P.const_tensors[i]];
    if (t.d.kind == YOTA_T_CONST) {
      const float v = (float)t.d.cval;
      YOTA_CUDA(cudaMemcpyAsync(P.arena + i * 256, &v, sizeof(float), cudaMemcpyHostToDevice, P.ctx->stream));
    }
  }
  YOTA_CUDA(cudaStreamSynchronize(P.ctx->stream));
  return YOTA_OK;
}

static int bind_region(yota_program& P, Region& r, RegionLaunch& L) {
  const int n_store = L.p.n_out - (int)r.sums.size();
  for (int q = 0; q < L.p.n_in; q++) {
    if (L.p.in[q].cls == CLS_GATHER) {  // x[:, idx[c]]: read from the gathered array itself
      L.p.in[q].ptr = static_cast<const float*>(tensor_ptr(P, r.gather_x[q]));
      L.p.in[q].idx = static_cast<const long long*>(tensor_ptr(P, r.gather_idx[q]));
      L.p.in[q].gather_n = P.T[r.gather_x[q]].d.dims[1];
      L.p.in[q].err = P.ctx->err_dev;
    } else
      L.p.in[q].ptr = static_cast<const float*>(tensor_ptr(P, r.in_tensors[q]));
    if (L.V == 4 && (reinterpret_cast<uintptr_t>(L.p.in[q].ptr) & 15) &&
        (L.p.in[q].cls == CLS_FULL || L.p.in[q].cls == CLS_ROWVEC || L.p.in[q].cls == CLS_GATHER))
      YOTA_FAIL(YOTA_E_ARG, "array bound to the program is not 16-byte aligned (use yota_alloc)");
  }
  for (int q = 0; q < L.p.n_out; q++) {
    if (q < n_store) {
      L.p.out[q].ptr = static_cast<float*>(tensor_ptr(P, r.out_tensors[q]));
      if (L.V == 4 && (reinterpret_cast<uintptr_t>(L.p.out[q].ptr) & 15))
        YOTA_FAIL(YOTA_E_ARG, "array bound to the program is not 16-byte aligned (use yota_alloc)");
    } else
      L.p.out[q].ptr = reinterpret_cast<float*>(P.arena + r.ws_off[q]);
  }
  return YOTA_OK;
}

static int run_gemm_epilogue(yota_program& P, Step& st, const GemmArgs& g_in, cudaStream_t s) {
  GemmArgs g = g_in;
  if (st.tsh_make >= 0) {
    const yota_program::Shadow& sh = P.tshadows[st.tsh_make];
    g.tsh = reinterpret_cast<float*>(P.arena + sh.arena_off);
    g.tsh_ld = P.T[sh.tensor].dims[1];
    g.tsh_out = st.tsh_make_out;
  }
  if (st.bsh_make >= 0) {
    g.bsh = P.arena + P.shadows[st.bsh_make].arena_off;
    g.bsh_out = st.bsh_make_out;
  }
  if (st.rs_finish >= 0) {
    const Step& fs = P.steps[st.rs_finish];
    g.rs_out = static_cast<float*>(tensor_ptr(P, fs.out_tensor));
    g.rs_scale = fs.scale;
    g.rs_cnt = reinterpret_cast<unsigned*>(P.arena + P.counter_off);
  }
  Region& r = P.regions[st.epi_region];
  RegionLaunch L = r.L;
  L.p.in[st.epi_z_in].ptr = nullptr;  // the accumulator stands in for this operand
  for (int q = 0; q < L.p.n_in; q++)
    if (q != st.epi_z_in) L.p.in[q].ptr = static_cast<const float*>(tensor_ptr(P, r.in_tensors[q]));
  const int n_store = L.p.n_out - (int)r.sums.size();
  for (int q = 0; q < L.p.n_out; q++)
    L.p.out[q].ptr = q < n_store ? static_cast<float*>(tensor_ptr(P, r.out_tensors[q]))
                                 : reinterpret_cast<float*>(P.arena + r.ws_off[q]);
  return launch_gemm_tc_epilogue(P.ctx, g, L.static_id, L.p, st.epi_z_in, s);
}

// GemmArgs and bound region parameters of a GEMM + epilogue step in plain TF32 mode (no shadows)
static void bind_epilogue_step(yota_program& P, const Step& st, GemmArgs& g, RegionLaunch& L) {
  const yota_op_desc& o = P.ops[st.op];
  g = GemmArgs{};
  g.transA = (int)o.iattr[0];
  g.transB = (int)o.iattr[1];
  gemm_dims(P, o, &g.M, &g.N, &g.K, &g.lda, &g.ldb);
  g.A = tensor_ptr(P, o.in[0]);
  g.B = tensor_ptr(P, o.in[1]);
  g.C = static_cast<float*>(tensor_ptr(P, o.out));
  g.ldc = g.M;
  g.mode = 1;
  Region& r = P.regions[st.epi_region];
  L = r.L;
  L.p.in[st.epi_z_in].ptr = nullptr;
  for (int q = 0; q < L.p.n_in; q++)
    if (q != st.epi_z_in) L.p.in[q].ptr = static_cast<const float*>(tensor_ptr(P, r.in_tensors[q]));
  for (int q = 0; q < L.p.n_out; q++) L.p.out[q].ptr = static_cast<float*>(tensor_ptr(P, r.out_tensors[q]));
}

// Head of a marked recurrence: verify on the bound addresses that the steps are one arithmetic
// progression (A fixed; B, the chain's FULL operands and its stores advancing by constant strides;
// B of step i + 1 = a store of step i; nothing a later step reads is written inside the run except
// through B) and launch them as one persistent kernel.  Returns YOTA_OK (launched), 1 (launch the
// steps one by one) or an error.
static int try_run_chain(yota_program& P, int head, cudaStream_t s) {
  const Step& hs = P.steps[head];
  std::vector<int> mem;
  for (size_t q = head; q < P.steps.size() && (int)mem.size() < hs.chain_len; q++)
    if (P.steps[q].chain_head == head) mem.push_back((int)q);
  const int n = (int)mem.size();
  if (n != hs.chain_len || n < 2) return 1;
  std::vector<GemmArgs> G(n);
  std::vector<RegionLaunch> R(n);
  for (int i = 0; i < n; i++) bind_epilogue_step(P, P.steps[mem[i]], G[i], R[i]);
  const GemmArgs& g0 = G[0];
  const RegionParams& r0 = R[0].p;
  const int z = hs.epi_z_in;
  const long long b_delta = static_cast<const float*>(G[1].B) - static_cast<const float*>(g0.B);
  long long e_delta = 0;
  bool have_e = false;
  auto full_ptr_delta = [&](int i, long long d) {  // every moving operand moves by the same stride
    if (!have_e) {
      e_delta = d / (i > 0 ? i : 1);
      have_e = true;
    }
    return d == e_delta * i;
  };
  for (int i = 1; i < n; i++) {
    const GemmArgs& g = G[i];
    const RegionParams& r = R[i].p;
    if (g.A != g0.A || g.lda != g0.lda || g.ldb != g0.ldb || g.M != g0.M || g.N != g0.N || g.K != g0.K) return 1;
    if (static_cast<const float*>(g.B) - static_cast<const float*>(g0.B) != b_delta * i) return 1;
    if (r.n_in != r0.n_in || r.n_out != r0.n_out || r.n_ops != r0.n_ops) return 1;
    for (int q = 0; q < r.n_ops; q++)
      if (memcmp(&r.ops[q], &r0.ops[q], sizeof(MicroOp)) != 0) return 1;
    for (int q = 0; q < r.n_in; q++) {
      if (q == z) continue;
      if (r.in[q].cls != r0.in[q].cls) return 1;
      if (r.in[q].cls == CLS_FULL) {
        if (!full_ptr_delta(i, r.in[q].ptr - r0.in[q].ptr)) return 1;
      } else if (r.in[q].ptr != r0.in[q].ptr)
        return 1;
    }
    for (int q = 0; q < r.n_out; q++)
      if (!full_ptr_delta(i, r.out[q].ptr - r0.out[q].ptr)) return 1;
    bool fed = false;  // B of this step is what the previous step stored
    for (int q = 0; q < r.n_out; q++) fed = fed || R[i - 1].p.out[q].ptr == static_cast<const float*>(g.B);
    if (!fed) return 1;
  }
  if (!have_e || e_delta % g0.M) return 1;
  // the chain's other operands are read without waiting for anything: none of them may live in
  // memory the run itself writes
  const long long tile = g0.M * g0.N, span = e_delta * (n - 1);
  for (int qo = 0; qo < r0.n_out; qo++) {
    const float* o_lo = r0.out[qo].ptr + (span < 0 ? span : 0);
    const float* o_hi = r0.out[qo].ptr + (span > 0 ? span : 0) + tile;
    for (int q = 0; q < r0.n_in; q++) {
      if (q == z) continue;
      const bool moves = r0.in[q].cls == CLS_FULL;
      const float* i_lo = r0.in[q].ptr + (moves && span < 0 ? span : 0);
      const float* i_hi = r0.in[q].ptr + (moves && span > 0 ? span : 0) + (moves ? tile : g0.M);
      if (i_lo < o_hi && o_lo < i_hi) return 1;
    }
    const float* a_lo = static_cast<const float*>(g0.A);
    if (a_lo < o_hi && o_lo < a_lo + g0.M * g0.K) return 1;
  }
  unsigned* flags = reinterpret_cast<unsigned*>(P.arena + P.chain_flags_off + (size_t)hs.chain_id * 256);
  if ((g0.M / 128) * (g0.N / 32) > 64) return 1;  // 256 bytes of flags per recurrence
  return launch_gemm_chain(P.ctx, g0, R[0].static_id, r0, z, n, b_delta, e_delta / g0.M, flags, s);
}

static int run_step(yota_program& P, Step& st, cudaStream_t s, int unget_phase = UNGET_ALL) {
  switch (st.kind) {
    case STEP_REGION: {
      Region& r = P.regions[st.region];
      RegionLaunch L = r.L;
      YOTA_TRY(bind_region(P, r, L));
      return launch_region(L, s);
    }
    case STEP_NOP: return YOTA_OK;
    case STEP_ROWSUM_FINISH:
      if (st.fused_away) return YOTA_OK;  // reduced inside the GEMM that produced the partials
      return launch_rowsum_finish(static_cast<float*>(tensor_ptr(P, st.out_tensor)),
                                  reinterpret_cast<const float*>(P.arena + st.ws_off), st.R, st.P, st.scale, s);
    case STEP_SCALAR_FINISH:
      return launch_scalar_finish(static_cast<float*>(tensor_ptr(P, st.out_tensor)),
                                  reinterpret_cast<const float*>(P.arena + st.ws_off), st.P, st.scale, s);
    default: break;
  }
  const yota_op_desc& o = P.ops[st.op];
  float* out = static_cast<float*>(tensor_ptr(P, st.out_override >= 0 ? st.out_override : o.out));
  const float* in0 = static_cast<const float*>(tensor_ptr(P, o.in[0]));
  const TInfo& ot = P.T[o.out];
  switch (st.kind) {
    case STEP_COLSUM: return launch_colsum(out, in0, st.R, st.C, st.scale, s);
    case STEP_GENERIC_SUM: {
      const TInfo& x = P.T[o.in[0]];
      GenericSum g{};
      g.nd = x.nd;
      for (int d = 0; d < 4; d++) g.dims[d] = x.dims[d];
      g.mask = (unsigned)o.iattr[0] & ((1u << x.nd) - 1);
      g.scale = st.scale;
      return launch_generic_sum(g, out, in0, s);
    }
    case STEP_GENERIC_EW: {
      GenericEw g{};
      g.opcode = o.opcode;
      g.nd = ot.nd;
      g.imm = (float)o.fattr[0];
      const TInfo& a = P.T[o.in[0]];
      const TInfo& b = P.T[o.in[o.n_in > 1 ? 1 : 0]];
      long long sa = 1, sb = 1;
      for (int d = 0; d < 4; d++) {
        g.dims[d] = ot.dims[d];
        g.sa[d] = (d < a.nd && a.dims[d] != 1) ? sa : 0;
        g.sb[d] = (d < b.nd && b.dims[d] != 1) ? sb : 0;
        sa *= d < a.nd ? a.dims[d] : 1;
        sb *= d < b.nd ? b.dims[d] : 1;
      }
      const float* in1 = o.n_in > 1 ? static_cast<const float*>(tensor_ptr(P, o.in[1])) : nullptr;
      return launch_generic_ew(g, out, in0, in1, s);
    }
    case STEP_GEMM: {
      GemmArgs g{};
      g.transA = (int)o.iattr[0];
      g.transB = (int)o.iattr[1];
      gemm_dims(P, o, &g.M, &g.N, &g.K, &g.lda, &g.ldb);
      g.A = in0;
      g.B = tensor_ptr(P, o.in[1]);
      g.C = out;
      g.ldc = g.M;
      g.accumulate = st.acc_tensor >= 0 ? 1 : 0;
      g.mode = gemm_mode_of(P.flags);
      g.pdl = st.pdl;
      g.early_a = st.early_a;
      if (st.ksplit > 1 && st.ws_off >= 0) {
        g.ksplit = st.ksplit;
        g.splitk_ws = P.arena + st.ws_off;
      }
      if ((g.mode == 2 || g.mode == 3) && st.shadow[0] >= 0) {
        void* sh[2];
        for (int j = 0; j < 2; j++) {
          char* base = P.arena + P.shadows[st.shadow[j]].arena_off;  // shadow of the whole root array
          if (st.cast[j]) {
            const int root = root_of(P, o.in[j]);
            const float* src = static_cast<const float*>(root_ptr(P, root));
            if (g.mode == 2)
              YOTA_TRY(launch_cast_bf16(base, src, P.T[root].numel, s));
            else
              YOTA_TRY(launch_split_tf32(reinterpret_cast<float*>(base), src, P.T[root].numel, s));
          }
          sh[j] = base + alias_offset(P, o.in[j]) * P.shadow_elem_bytes;
        }
        if (g.mode == 2) {
          g.A = sh[0];
          g.B = sh[1];
        } else {
          g.A_lo = sh[0];
          g.B_lo = sh[1];
        }
        if (st.epi_region >= 0) return run_gemm_epilogue(P, st, g, s);
        return launch_gemm_tc(P.ctx, g, s);
      }
      if (g.mode == 2 || g.mode == 3) g.mode = 0;  // not tile-aligned: no shadows were planned, strict-FP32 kernel
      if (st.epi_region >= 0) return run_gemm_epilogue(P, st, g, s);
      if (g.mode == 1) {  // operands with a transposed copy are read K-major from it
        if (st.tsh_use[0] >= 0) {
          g.A = P.arena + P.tshadows[st.tsh_use[0]].arena_off;
          g.transA = 1;
          g.lda = g.K;
        }
        if (st.tsh_use[1] >= 0) {
          g.B = P.arena + P.tshadows[st.tsh_use[1]].arena_off;
          g.transB = 0;
          g.ldb = g.K;
        }
      }
      if (g.mode == 1 && gemm_tc_eligible(g)) return launch_gemm_tc(P.ctx, g, s);
      g.mode = 0;
      return launch_gemm_ffma(g, s);
    }
    case STEP_SLICE_ACCUM: {
      IndexSpec sp;
      long long nsel;
      YOTA_TRY(index_spec_of(P, o, ot, sp, &nsel));
      return launch_slice_accum(sp, out, in0, P.T[o.in[0]].numel == 1, s);
    }
    case STEP_GETINDEX:
    case STEP_UNGETINDEX: {
      IndexSpec sp;
      long long nsel;
      const TInfo& x = st.kind == STEP_GETINDEX ? P.T[o.in[0]] : ot;
      YOTA_TRY(index_spec_of(P, o, x, sp, &nsel));
      for (int d = 0; d < sp.n; d++)
        if (sp.kind[d] == YOTA_IDX_VEC)
          sp.vec[d] = static_cast<const int64_t*>(tensor_ptr(P, o.in[1 + (int)sp.a[d]]));
      if (st.kind == STEP_GETINDEX) return launch_getindex(sp, out, in0, ot.numel, s, P.ctx->err_dev);
      const float* scale =
          P.op_scale[st.op] >= 0 ? static_cast<const float*>(tensor_ptr(P, P.op_scale[st.op])) : nullptr;
      return launch_ungetindex(sp, out, in0, st.ws_off >= 0 ? P.arena + st.ws_off : nullptr, st.ws_bytes,
                               P.ctx->sm_count, s, scale, P.ctx->err_dev, unget_phase);
    }
    case STEP_LOGSOFTMAX: {
      const TInfo& x = P.T[o.in[0]];
      const long long R = x.nd >= 1 ? x.dims[0] : 1;
      return launch_logsoftmax(out, in0, R, x.numel / std::max(1ll, R), s);
    }
    case STEP_TRANSPOSE: {
      const TInfo& x = P.T[o.in[0]];
      const long long R = x.d.ndim >= 1 ? x.d.dims[0] : 1, C = x.d.ndim == 2 ? x.d.dims[1] : 1;
      if (R == 1 || C == 1) {  // vector <-> row matrix: same memory
        YOTA_CUDA(cudaMemcpyAsync(out, in0, (size_t)x.numel * 4, cudaMemcpyDeviceToDevice, s));
        return YOTA_OK;
      }
      return launch_transpose(out, in0, R, C, s);
    }
    case STEP_CAT: {
      const int dim = (int)o.iattr[0] - 1;
      long long odims[4], ostr[4];
      for (int d = 0; d < 4; d++) odims[d] = d < ot.d.ndim ? ot.d.dims[d] : 1;
      ostr[0] = 1;
      for (int d = 1; d < 4; d++) ostr[d] = ostr[d - 1] * odims[d - 1];
      long long off = 0;
      for (int j = 0; j < o.n_in; j++) {
        const TInfo& x = P.T[o.in[j]];
        long long xd[4], xs[4];
        for (int d = 0; d < 4; d++) xd[d] = d < x.d.ndim ? x.d.dims[d] : 1;
        xs[0] = 1;
        for (int d = 1; d < 4; d++) xs[d] = xs[d - 1] * xd[d - 1];
        YOTA_TRY(launch_copy_strided(out + off * ostr[dim], static_cast<const float*>(tensor_ptr(P, o.in[j])), 4, xd,
                                     ostr, xs, s));
        off += xd[dim];
      }
      return YOTA_OK;
    }
    case STEP_SLICE: {
      const TInfo& x = P.T[o.in[0]];
      const int dim = (int)o.iattr[0] - 1;
      long long xd[4], xs[4], od[4], os[4];
      for (int d = 0; d < 4; d++) xd[d] = d < x.d.ndim ? x.d.dims[d] : 1;
      for (int d = 0; d < 4; d++) od[d] = xd[d];
      od[dim] = o.iattr[2] - o.iattr[1] + 1;
      xs[0] = os[0] = 1;
      for (int d = 1; d < 4; d++) {
        xs[d] = xs[d - 1] * xd[d - 1];
        os[d] = os[d - 1] * od[d - 1];
      }
      return launch_copy_strided(out, in0 + (o.iattr[1] - 1) * xs[dim], 4, od, os, xs, s);
    }
    default: break;
  }
  YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "internal: unknown step kind %d", st.kind);
}

// optional timeline of one eager run: device start/end of every step (compute stream) and of
// every all-reduce bucket (comm stream), plus the host time at which each was enqueued
struct Timeline {
  struct Rec {
    std::string label;
    cudaEvent_t e0, e1;
    double host_us;
  };
  std::vector<Rec> recs;
  cudaEvent_t base = nullptr;
  std::chrono::steady_clock::time_point t0;
  int open(const std::string& label, cudaStream_t s) {
    Rec r;
    r.label = label;
    YOTA_CUDA(cudaEventCreate(&r.e0));
    YOTA_CUDA(cudaEventCreate(&r.e1));
    r.host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    YOTA_CUDA(cudaEventRecord(r.e0, s));
    recs.push_back(r);
    return YOTA_OK;
  }
  int close(cudaStream_t s) {
    YOTA_CUDA(cudaEventRecord(recs.back().e1, s));
    return YOTA_OK;
  }
};

// out slot whose caller-owned buffer tensor t lives in (-1: program-owned or an input)
static int out_slot_of(const yota_program& P, int t) {
  const TInfo& r = P.T[root_of(P, t)];
  return r.d.kind == YOTA_T_OUTPUT ? r.d.slot : r.fwd_out_slot;
}

// Which step first / last touches every output buffer.  A reduction is issued after the LAST
// step that reads or writes the buffer (a later step may still read an output: two parameters
// sharing one tangent, an output consumed by later ops); the first touch is where a run must
// wait for a reduction of the same buffer left in flight by the previous run.
static int plan_comm(yota_program& P) {
  yota_program::CommPlan& C = P.comm;
  if (C.valid) return YOTA_OK;
  const size_t ns = P.steps.size();
  C.out_bytes.assign(P.n_out, 0);
  C.in_bytes.assign(P.n_in, 0);
  for (auto& t : P.T) {
    const long long b = t.numel * (t.d.dtype == YOTA_I64 ? 8 : 4);
    if (t.d.kind == YOTA_T_OUTPUT && t.d.slot >= 0 && t.d.slot < P.n_out) C.out_bytes[t.d.slot] = std::max(C.out_bytes[t.d.slot], b);
    if (t.d.kind == YOTA_T_INPUT && t.d.slot >= 0 && t.d.slot < P.n_in) C.in_bytes[t.d.slot] = std::max(C.in_bytes[t.d.slot], b);
  }
  std::vector<int> first(P.n_out, -1), last(P.n_out, -1);
  auto touch = [&](size_t q, int t) {
    const int slot = out_slot_of(P, t);
    if (slot < 0 || slot >= P.n_out) return;
    if (first[slot] < 0) first[slot] = (int)q;
    last[slot] = (int)q;
  };
  for (size_t q = 0; q < ns; q++) {
    const Step& st = P.steps[q];
    if (st.kind == STEP_NOP) continue;
    for (int t : st.reads) touch(q, t);
    for (int t : st.writes) touch(q, t);
    if (st.out_tensor >= 0) touch(q, st.out_tensor);
    if (st.op >= 0 && st.kind != STEP_ROWSUM_FINISH && st.kind != STEP_SCALAR_FINISH) {
      const yota_op_desc& o = P.ops[st.op];
      for (int j = 0; j < o.n_in; j++) touch(q, o.in[j]);
      touch(q, st.out_override >= 0 ? st.out_override : o.out);
    }
  }
  C.first_touch_at.assign(ns, {});
  C.ar_at.assign(ns, {});
  for (int slot = 0; slot < P.n_out; slot++)
    if (first[slot] >= 0) C.first_touch_at[first[slot]].push_back(slot);
  C.last_ar_step = 0;
  for (int slot : P.allreduce_slots) {
    if (last[slot] < 0)
      YOTA_FAIL(YOTA_E_ARG, "all-reduce marked for output slot %d, which no step of the program writes", slot);
    C.ar_at[last[slot]].push_back(slot);
    C.last_ar_step = std::max(C.last_ar_step, (size_t)last[slot]);
  }
  C.valid = true;
  return YOTA_OK;
}

static int env_int_p(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static int run_all_steps(yota_program& P, cudaStream_t s, bool with_comm, Timeline* tl = nullptr) {
  yota_ctx* ctx = P.ctx;
  struct ReserveGuard {  // the SM reservation never outlives this sweep, error paths included
    yota_ctx* c;
    ~ReserveGuard() { c->sm_reserved = 0; }
  } guard{ctx};
  ctx->sm_reserved = 0;
  YOTA_TRY(plan_comm(P));
  const yota_program::CommPlan& C = P.comm;
  // Reductions of an earlier run may still be in flight (the previous eval's tail buckets overlap
  // this forward sweep): order this run behind the ones that touch its buffers, and leave their
  // CTAs some SMs during the first GEMMs.
  const bool inherited = !ctx->pending_ar.empty();
  if (inherited)
    for (int i = 0; i < P.n_in; i++) YOTA_TRY(ctx_wait_pending(ctx, P.bound_in[i], (size_t)C.in_bytes[i], s));
  // Sorts that need the index vectors only start now, beside the forward sweep (fork / join by
  // events: inside a stream capture this becomes a parallel branch of the graph).
  if (P.n_presort > 0) {
    if (!P.presort_fork) YOTA_CUDA(cudaEventCreateWithFlags(&P.presort_fork, cudaEventDisableTiming));
    while ((int)P.presort_join.size() < P.n_presort) {
      cudaEvent_t e;
      YOTA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      P.presort_join.push_back(e);
    }
    YOTA_CUDA(cudaEventRecord(P.presort_fork, s));
    YOTA_CUDA(cudaStreamWaitEvent(ctx->side_stream, P.presort_fork, 0));
    for (Step& st : P.steps)
      if (st.presort >= 0) {
        if (tl) YOTA_TRY(tl->open("index sort for " + st.label, ctx->side_stream));
        YOTA_TRY(run_step(P, st, ctx->side_stream, UNGET_SORT));
        if (tl) YOTA_TRY(tl->close(ctx->side_stream));
        YOTA_CUDA(cudaEventRecord(P.presort_join[st.presort], ctx->side_stream));
      }
  }
  int head_gemms = inherited && with_comm ? env_int_p("YOTA_AR_HEAD_RESERVE", 3) : 0;
  const int reserve = with_comm ? comm_reserved_sms(ctx) : 0;
  bool in_flight = false;
  int chain_done_until = -1;  // steps of a recurrence up to here were executed by their head's persistent launch
  std::vector<int> forked;    // leaves running on side streams
  size_t ev = 0;
  // Buckets: outputs finished by a step wait until at least kBucketBytes are pending (a
  // layer's 16 KiB bias gradient rides with the next weight gradient instead of paying a
  // launch and two cross-GPU barriers of its own); the last step flushes the rest.
  constexpr long long kBucketBytes = 1 << 20;
  std::vector<float*> pend_buf;
  std::vector<long long> pend_n;
  long long pend_bytes = 0;
  for (size_t q = 0; q < P.steps.size(); q++) {
    Step& st = P.steps[q];
    if (!ctx->pending_ar.empty())
      for (int slot : C.first_touch_at[q])
        YOTA_TRY(ctx_wait_pending(ctx, P.bound_out[slot], (size_t)C.out_bytes[slot], s));
    if (st.kind == STEP_GEMM) {
      ctx->sm_reserved = (in_flight || head_gemms > 0) ? reserve : 0;
      if (head_gemms > 0) head_gemms--;
    }
    const bool launches = st.kind != STEP_NOP && !st.fused_away;
    const bool leaf = st.fork >= 0 && !with_comm;
    if (tl && launches && !leaf) YOTA_TRY(tl->open(st.label, s));
    if (st.presort >= 0) YOTA_CUDA(cudaStreamWaitEvent(s, P.presort_join[st.presort], 0));
    if (leaf) {  // small leaf: beside the sweep, joined after the last step
      while ((int)P.fork_ev.size() < P.n_forks) {
        cudaEvent_t a, b;
        YOTA_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        YOTA_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        P.fork_ev.push_back(a);
        P.join_ev.push_back(b);
      }
      cudaStream_t lane = ctx->leaf_streams[st.fork % 3];
      YOTA_CUDA(cudaEventRecord(P.fork_ev[st.fork], s));
      YOTA_CUDA(cudaStreamWaitEvent(lane, P.fork_ev[st.fork], 0));
      if (tl && launches) YOTA_TRY(tl->open(st.label, lane));
      YOTA_TRY(run_step(P, st, lane));
      if (tl && launches) YOTA_TRY(tl->close(lane));
      YOTA_CUDA(cudaEventRecord(P.join_ev[st.fork], lane));
      forked.push_back(st.fork);
      continue;
    }
    if (st.chain_len > 0) {  // head of a recurrence: one persistent launch for all its steps, if they line up
      // (not when a member is where a bound output is first touched or all-reduced: those waits and
      // launches are tied to the member's own place in the sweep)
      bool plain = true;
      int last = -1, seen = 0;
      for (size_t e = q; e < P.steps.size() && seen < st.chain_len; e++)
        if (P.steps[e].chain_head == (int)q) {
          seen++;
          last = (int)e;
          plain = plain && C.first_touch_at[e].empty() && C.ar_at[e].empty();
        }
      const int taken = plain ? try_run_chain(P, (int)q, s) : 1;
      if (taken < 0) return taken;
      chain_done_until = taken == YOTA_OK ? last : -1;
      P.chain_taken[st.chain_id] = taken == YOTA_OK;
    }
    if (!(st.chain_head >= 0 && (int)q <= chain_done_until))
      YOTA_TRY(run_step(P, st, s, st.presort >= 0 ? UNGET_FOLD : UNGET_ALL));
    if (tl && launches) YOTA_TRY(tl->close(s));
    if (with_comm && !C.ar_at[q].empty()) {
      for (int slot : C.ar_at[q]) {
        pend_buf.push_back(static_cast<float*>(P.bound_out[slot]));
        pend_n.push_back(C.out_bytes[slot] / 4);
        pend_bytes += C.out_bytes[slot];
      }
      if (pend_bytes < kBucketBytes && q != C.last_ar_step) continue;
      // all-reduce the bucket on the comm stream while the compute stream continues with the
      // rest of the pullback sweep -- and, for the last buckets, with the next eval
      if (ev >= P.ar_events.size()) {
        cudaEvent_t e;
        YOTA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        P.ar_events.push_back(e);
      }
      YOTA_CUDA(cudaEventRecord(P.ar_events[ev], s));
      YOTA_CUDA(cudaStreamWaitEvent(ctx->comm_stream, P.ar_events[ev], 0));
      ev++;
      if (tl)
        YOTA_TRY(tl->open("allreduce " + std::to_string(pend_bytes) + " B (" + std::to_string(pend_buf.size()) +
                              " buffers) after step " + std::to_string(q),
                          ctx->comm_stream));
      YOTA_TRY(comm_allreduce_bucket(ctx, pend_buf.data(), pend_n.data(), (int)pend_buf.size(), ctx->comm_stream));
      if (tl) YOTA_TRY(tl->close(ctx->comm_stream));
      std::vector<std::pair<const char*, const char*>> ranges;
      for (size_t b = 0; b < pend_buf.size(); b++)
        ranges.push_back({reinterpret_cast<const char*>(pend_buf[b]), reinterpret_cast<const char*>(pend_buf[b] + pend_n[b])});
      YOTA_TRY(ctx_add_pending(ctx, ranges));
      pend_buf.clear();
      pend_n.clear();
      pend_bytes = 0;
      in_flight = true;  // from here on reductions run beside the sweep: leave their CTAs some SMs
    }
  }
  ctx->sm_reserved = 0;
  for (int k : forked) YOTA_CUDA(cudaStreamWaitEvent(s, P.join_ev[k], 0));
  if (with_comm && getenv("YOTA_AR_JOIN")) YOTA_TRY(ctx_join_comm(ctx, s));  // A/B: the round-1 behaviour
  // update!(param, grad) for the marked pairs: after every reader of the parameter in this sweep
  // and after its gradient has been summed over ranks (src/update.jl:42-49, x .= x .- lr .* gx)
  for (auto& u : P.updates) {
    const long long n = C.out_bytes[u.second] / 4;
    YOTA_TRY(ctx_wait_pending(ctx, P.bound_out[u.second], (size_t)n * 4, s));
    if (tl) YOTA_TRY(tl->open("update! input " + std::to_string(u.first), s));
    float* x = static_cast<float*>(P.bound_in[u.first]);
    const float* gx = static_cast<const float*>(P.bound_out[u.second]);
    if (P.adam) {
      const size_t k = (size_t)(&u - P.updates.data());
      YOTA_TRY(launch_adam(x, gx, P.adam_m[k], P.adam_v[k], n, P.update_lr, P.adam_b1, P.adam_b2, P.adam_eps, P.adam_t, 0, s));
    } else
      YOTA_TRY(launch_axpy(x, gx, n, P.update_lr, s));
    if (tl) YOTA_TRY(tl->close(s));
  }
  if (P.adam && !P.updates.empty()) YOTA_TRY(launch_tick(P.adam_t, s));  // one optimiser step per run
  return YOTA_OK;
}

static int bind(yota_program& P, void* const* in_ptrs, int64_t n_in, void* const* out_ptrs, int64_t n_out,
                bool* changed) {
  if (n_in < P.n_in || n_out < P.n_out)
    YOTA_FAIL(YOTA_E_ARG, "yota_program_run: program needs %d inputs / %d outputs", P.n_in, P.n_out);
  *changed = (int)P.bound_in.size() != P.n_in || (int)P.bound_out.size() != P.n_out;
  P.bound_in.resize(P.n_in);
  P.bound_out.resize(P.n_out);
  for (int i = 0; i < P.n_in; i++) {
    if (!in_ptrs[i]) YOTA_FAIL(YOTA_E_ARG, "yota_program_run: input %d is null", i);
    if (P.bound_in[i] != in_ptrs[i]) *changed = true;
    P.bound_in[i] = in_ptrs[i];
  }
  for (int i = 0; i < P.n_out; i++) {
    if (!out_ptrs[i]) YOTA_FAIL(YOTA_E_ARG, "yota_program_run: output %d is null", i);
    if (P.bound_out[i] != out_ptrs[i]) *changed = true;
    P.bound_out[i] = out_ptrs[i];
  }
  return YOTA_OK;
}

static int set_seed(yota_program& P, float seed) {
  if (P.seed_valid && P.last_seed == seed) return YOTA_OK;
  for (size_t i = 0; i < P.const_tensors.size(); i++)
    if (P.T[P.const_tensors[i]].d.kind == YOTA_T_SEED)
      YOTA_TRY(launch_fill_const(reinterpret_cast<float*>(P.arena + i * 256), seed, P.ctx->stream));
  P.last_seed = seed;
  P.seed_valid = true;
  return YOTA_OK;
}

}  // namespace yota

extern "C" int yota_program_run(yota_program* p, void* const* in_ptrs, int64_t n_in, void* const* out_ptrs,
                                int64_t n_out, float seed) {
  if (!p) YOTA_FAIL(YOTA_E_ARG, "null program");
  yota_program& P = *p;
  YOTA_TRY(ensure_arena(P));
  bool changed = false;
  YOTA_TRY(bind(P, in_ptrs, n_in, out_ptrs, n_out, &changed));
  YOTA_TRY(set_seed(P, seed));
  cudaStream_t s = P.ctx->stream;
  const bool with_comm = !P.allreduce_slots.empty() && P.ctx->nccl_comm;
  // With reductions in the sweep the run is issued eagerly: a captured graph would have to join
  // the communication stream before it ends, and the next launch would then wait for this eval's
  // last reductions instead of overlapping them with its forward sweep (run_all_steps).
  const bool use_graph = !(P.flags & YOTA_FLAG_NO_GRAPH) && !with_comm;
  (void)changed;
  if (!use_graph) return run_all_steps(P, s, with_comm);
  YOTA_TRY(ctx_join_comm(P.ctx, s));  // a replayed graph carries no waits for reductions left in flight
  // one graph per pointer binding: eager on the first run with a binding (also the warm-up
  // that sets function attributes), captured on the second, replayed afterwards
  yota_program::GraphEntry* ge = nullptr;
  for (auto& g : P.graphs)
    if (g.in == P.bound_in && g.out == P.bound_out) ge = &g;
  if (!ge) {
    if (P.graphs.size() >= 8) {  // evict the least recently used binding
      size_t lru = 0;
      for (size_t i = 1; i < P.graphs.size(); i++)
        if (P.graphs[i].last_use < P.graphs[lru].last_use) lru = i;
      if (P.graphs[lru].exec) cudaGraphExecDestroy(P.graphs[lru].exec);
      P.graphs.erase(P.graphs.begin() + lru);
    }
    P.graphs.emplace_back();
    ge = &P.graphs.back();
    ge->in = P.bound_in;
    ge->out = P.bound_out;
  }
  ge->last_use = ++P.run_counter;
  if (++ge->runs == 1) return run_all_steps(P, s, with_comm);
  if (!ge->exec) {
    cudaGraph_t g = nullptr;
    YOTA_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int st = run_all_steps(P, s, with_comm);
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (st != YOTA_OK) {
      if (g) cudaGraphDestroy(g);
      return st;
    }
    if (e != cudaSuccess) YOTA_FAIL(YOTA_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ge->exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      ge->exec = nullptr;
      YOTA_FAIL(YOTA_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    }
  }
  YOTA_CUDA(cudaGraphLaunch(ge->exec, s));
  return YOTA_OK;
}

extern "C" int yota_program_profile(yota_program* p, void* const* in_ptrs, int64_t n_in, void* const* out_ptrs,
                                    int64_t n_out, float seed, float* step_ms, int64_t cap_steps, char* names,
                                    int64_t cap_names) {
  if (!p) YOTA_FAIL(YOTA_E_ARG, "null program");
  yota_program& P = *p;
  YOTA_TRY(ensure_arena(P));
  bool changed = false;
  YOTA_TRY(bind(P, in_ptrs, n_in, out_ptrs, n_out, &changed));
  YOTA_TRY(set_seed(P, seed));
  cudaStream_t s = P.ctx->stream;
  YOTA_TRY(ctx_join_comm(P.ctx, s));
  const size_t n = P.steps.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) YOTA_CUDA(cudaEventCreate(&e));
  YOTA_CUDA(cudaEventRecord(ev[0], s));
  for (size_t q = 0; q < n; q++) {
    YOTA_TRY(run_step(P, P.steps[q], s));
    YOTA_CUDA(cudaEventRecord(ev[q + 1], s));
  }
  YOTA_CUDA(cudaStreamSynchronize(s));
  std::string nm;
  for (size_t q = 0; q < n; q++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[q], ev[q + 1]);
    if ((int64_t)q < cap_steps && step_ms) step_ms[q] = ms;
    nm += P.steps[q].label;
    if (q + 1 < n) nm += "\n";
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (names && cap_names > 0) {
    const size_t m = std::min((size_t)cap_names - 1, nm.size());
    memcpy(names, nm.data(), m);
    names[m] = 0;
  }
  return (int)std::min<int64_t>((int64_t)n, cap_steps);
}

// Timeline of `reps` back-to-back eager runs (the last one is reported): for every launched
// step and every all-reduce bucket its device start / end (ms since the run's first event) and
// the host enqueue time (ms since the run began).  out: 3 floats per record.
extern "C" int yota_program_timeline(yota_program* p, void* const* in_ptrs, int64_t n_in, void* const* out_ptrs,
                                     int64_t n_out, float seed, int reps, float* rec_ms, int64_t cap_recs, char* names,
                                     int64_t cap_names) {
  if (!p) YOTA_FAIL(YOTA_E_ARG, "null program");
  yota_program& P = *p;
  YOTA_TRY(ensure_arena(P));
  bool changed = false;
  YOTA_TRY(bind(P, in_ptrs, n_in, out_ptrs, n_out, &changed));
  YOTA_TRY(set_seed(P, seed));
  cudaStream_t s = P.ctx->stream;
  const bool with_comm = !P.allreduce_slots.empty() && P.ctx->nccl_comm;
  for (int r = 0; r + 1 < reps; r++) YOTA_TRY(run_all_steps(P, s, with_comm));
  Timeline tl;
  YOTA_CUDA(cudaEventCreate(&tl.base));
  tl.t0 = std::chrono::steady_clock::now();
  YOTA_CUDA(cudaEventRecord(tl.base, s));
  const int st = run_all_steps(P, s, with_comm, &tl);
  cudaEvent_t end;
  YOTA_CUDA(cudaEventCreate(&end));
  YOTA_CUDA(cudaEventRecord(end, s));
  YOTA_CUDA(cudaStreamSynchronize(s));
  YOTA_CUDA(cudaStreamSynchronize(P.ctx->comm_stream));
  YOTA_TRY(ctx_join_comm(P.ctx, s));
  std::string nm;
  int64_t n = 0;
  for (auto& r : tl.recs) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, tl.base, r.e0);
    cudaEventElapsedTime(&b, tl.base, r.e1);
    if (n < cap_recs && rec_ms) {
      rec_ms[3 * n] = a;
      rec_ms[3 * n + 1] = b;
      rec_ms[3 * n + 2] = (float)(r.host_us * 1e-3);
    }
    nm += r.label + "\n";
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    n++;
  }
  float total = 0.f;
  cudaEventElapsedTime(&total, tl.base, end);
  if (n < cap_recs && rec_ms) {
    rec_ms[3 * n] = 0.f;
    rec_ms[3 * n + 1] = total;
    rec_ms[3 * n + 2] = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl.t0).count();
  }
  nm += "(whole run)";
  n++;
  cudaEventDestroy(tl.base);
  cudaEventDestroy(end);
  if (names && cap_names > 0) {
    const size_t m = std::min((size_t)cap_names - 1, nm.size());
    memcpy(names, nm.data(), m);
    names[m] = 0;
  }
  if (st != YOTA_OK) return st;
  return (int)std::min<int64_t>(n, cap_recs);
}

extern "C" int yota_program_set_update_adam(yota_program* p, const int32_t* in_slots, const int32_t* out_slots,
                                            int64_t n, float lr, float beta1, float beta2, float eps) {
  YOTA_TRY(yota_program_set_update(p, in_slots, out_slots, n, lr));  // validates the pairs, clears graphs and old state
  if (n == 0) return YOTA_OK;
  YOTA_TRY(ctx_ensure_device(p->ctx));
  YOTA_TRY(plan_comm(*p));
  for (auto& u : p->updates) {
    const size_t bytes = (size_t)p->comm.out_bytes[u.second];
    float *m = nullptr, *v = nullptr;
    if (cudaMalloc((void**)&m, bytes ? bytes : 4) != cudaSuccess || cudaMalloc((void**)&v, bytes ? bytes : 4) != cudaSuccess) {
      if (m) cudaFree(m);
      free_adam_state(p);
      YOTA_FAIL(YOTA_E_NOMEM, "yota_program_set_update_adam: cannot allocate the moment arrays");
    }
    YOTA_CUDA(cudaMemsetAsync(m, 0, bytes, p->ctx->stream));
    YOTA_CUDA(cudaMemsetAsync(v, 0, bytes, p->ctx->stream));
    p->adam_m.push_back(m);
    p->adam_v.push_back(v);
  }
  YOTA_CUDA(cudaMalloc((void**)&p->adam_t, sizeof(int)));
  YOTA_CUDA(cudaMemsetAsync(p->adam_t, 0, sizeof(int), p->ctx->stream));
  p->adam = true;
  p->adam_b1 = beta1;
  p->adam_b2 = beta2;
  p->adam_eps = eps;
  return YOTA_OK;
}

extern "C" int yota_program_set_update(yota_program* p, const int32_t* in_slots, const int32_t* out_slots, int64_t n,
                                       float lr) {
  if (!p || n < 0 || (n > 0 && (!in_slots || !out_slots))) YOTA_FAIL(YOTA_E_ARG, "yota_program_set_update: bad arguments");
  std::vector<std::pair<int, int>> ups;
  for (int64_t i = 0; i < n; i++) {
    const int a = in_slots[i], b = out_slots[i];
    if (a < 0 || a >= p->n_in || b < 0 || b >= p->n_out) YOTA_FAIL(YOTA_E_ARG, "yota_program_set_update: bad slot");
    long long na = -1, nb = -2;
    for (auto& t : p->T) {
      if (t.d.kind == YOTA_T_INPUT && t.d.slot == a && t.d.dtype == YOTA_F32) na = t.numel;
      if (t.d.kind == YOTA_T_OUTPUT && t.d.slot == b) nb = t.numel;
    }
    if (na != nb)
      YOTA_FAIL(YOTA_E_SHAPE, "yota_program_set_update: input %d and output %d are not Float32 arrays of one size", a, b);
    ups.push_back({a, b});
  }
  free_adam_state(p);
  p->updates = ups;
  p->update_lr = lr;
  for (auto& g : p->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  p->graphs.clear();
  return YOTA_OK;
}

extern "C" int yota_program_set_allreduce(yota_program* p, const int32_t* out_slots, int64_t n) {
  if (!p) YOTA_FAIL(YOTA_E_ARG, "null program");
  p->allreduce_slots.assign(out_slots, out_slots + n);
  p->comm.valid = false;
  for (int s : p->allreduce_slots)
    if (s < 0 || s >= p->n_out) YOTA_FAIL(YOTA_E_ARG, "yota_program_set_allreduce: bad out slot %d", s);
  if (plan_comm(*p) != YOTA_OK) {  // an all-reduced slot that no step writes
    p->allreduce_slots.clear();
    p->comm.valid = false;
    return YOTA_E_ARG;
  }
  for (auto& g : p->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  p->graphs.clear();
  return YOTA_OK;
}
