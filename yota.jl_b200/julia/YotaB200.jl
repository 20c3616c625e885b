# YotaB200.jl -- Julia glue between Yota.jl and libyota_cuda.so.
#
# STATUS: UNTESTED.  No Julia toolchain exists in the build image (SURVEY.md §8(c)), so this
# file has never been executed.  It is written against include/yota_cuda.h and mirrors, line
# for line, the Python host mirror that IS tested (yota.jl_b200/grad.py `lower`,
# yota.jl_b200/rules.py, yota.jl_b200/ffi.py).  It is the first thing to run once a Julia box
# exists (SURVEY.md §8(f) rank 2).
#
# What stays Yota's own (unchanged): Umlaut tracing, `gradtape`, `chainrules_transform!`,
# `back!`, `finalize_grad!`, ChainRules `rrule` dispatch (src/grad.jl:134-289).
# What this file replaces: `grad_compile(tape)` (src/grad.jl:293-307) for tapes whose array
# arguments are `B200Array`s -- the finished `Tape{GradCtx}` is lowered to the flat op program
# of include/yota_cuda.h and executed by `yota_program_run`.
module YotaB200

using Yota
using Yota: GradCtx, YotaRuleConfig, YOTA_RULE_CONFIG, _getfield
using Umlaut
using Umlaut: Tape, Variable, V, Call, Constant, Input, inputs
using ChainRulesCore: rrule, unthunk, ZeroTangent, NoTangent
import Statistics, NNlib
import Base.Broadcast: broadcasted, materialize

const LIB = joinpath(@__DIR__, "..", "lib", "libyota_cuda.so")

# ---------------------------------------------------------------------------------------
# C ABI structs (include/yota_cuda.h) -- isbits, same field order and padding
# ---------------------------------------------------------------------------------------
struct TensorDesc
    dtype::Int32
    ndim::Int32
    dims::NTuple{4,Int64}
    kind::Int32
    slot::Int32
    cval::Float64
end

struct OpDesc
    opcode::Int32
    n_in::Int32
    in::NTuple{6,Int32}
    out::Int32
    _pad::Int32
    iattr::NTuple{16,Int64}
    fattr::NTuple{2,Float64}
end

const F32, I64 = Int32(0), Int32(1)
const T_INPUT, T_TEMP, T_OUTPUT, T_CONST, T_SEED = Int32.(0:4)
const EW = (copy=1, add=2, sub=3, mul=4, div=5, neg=6, tanh=7, exp=8, log=9, sin=10, cos=11, sqrt=12,
            relu=13, sigmoid=14, square=15, gt0=16, addc=17, mulc=18, rsubc=19, rdivc=20, powc=21)
const OP = (sum=64, matmul=65, getindex=66, ungetindex=67, logsoftmax=68, reshape=69, transpose=70, cat=71, slice=72)
const IDX_COLON, IDX_INT, IDX_RANGE, IDX_VEC = 0, 1, 2, 3
# yota_program_build flags (include/yota_cuda.h; tests/test_abi.py compares the values with the header)
const FLAG_NO_FUSE = UInt32(0x1)
const FLAG_NO_GRAPH = UInt32(0x2)
const FLAG_GEMM_TF32 = UInt32(0x10)
const FLAG_GEMM_BF16 = UInt32(0x20)
const FLAG_NO_STATIC = UInt32(0x40)
const FLAG_GEMM_TF32X3 = UInt32(0x80)

check(status::Integer) = status == 0 ? nothing :
    error("libyota_cuda: " * unsafe_string(ccall((:yota_last_error, LIB), Cstring, ())) * " (status $status)")

# ---------------------------------------------------------------------------------------
# context and device arrays (the CuArray of this backend; src dispatch as in
# test/test_grad.jl:262-267)
# ---------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
function context()
    if CTX[] === nothing
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:yota_init, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), parse(Int, get(ENV, "LOCAL_RANK", "0")), h))
        CTX[] = Context(h[])
    end
    return CTX[]
end

mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    function B200Array{T,N}(dims::NTuple{N,Int}) where {T,N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:yota_alloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}),
                    context().h, max(prod(dims) * sizeof(T), 4), p))
        a = new{T,N}(p[], dims)
        finalizer(x -> ccall((:yota_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), context().h, x.ptr), a)
        return a
    end
    # non-owning view of memory allocated elsewhere (symmetric memory: symmetric_array)
    B200Array{T,N}(ptr::Ptr{Cvoid}, dims::NTuple{N,Int}) where {T,N} = new{T,N}(ptr, dims)
end
Base.size(a::B200Array) = a.dims
Base.getindex(::B200Array, i...) = error("scalar indexing of a B200Array is disabled; use Array(x)")

"Move a host array to the device (the `cu` of this backend)."
function b200(x::Array{T,N}) where {T<:Union{Float32,Int64},N}
    a = B200Array{T,N}(size(x))
    check(ccall((:yota_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context().h, a.ptr, x, sizeof(x)))
    return a
end
b200(x::AbstractArray{<:AbstractFloat}) = b200(Array{Float32}(x))
function Base.Array(a::B200Array{T,N}) where {T,N}
    x = Array{T,N}(undef, a.dims)
    check(ccall((:yota_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context().h, x, a.ptr, sizeof(x)))
    return x
end

# ---------------------------------------------------------------------------------------
# op-program builder
# ---------------------------------------------------------------------------------------
mutable struct Program
    tensors::Vector{TensorDesc}
    ops::Vector{OpDesc}
    n_in::Int
    n_out::Int
end
Program() = Program(TensorDesc[], OpDesc[], 0, 0)

pad4(d) = ntuple(i -> i <= length(d) ? Int64(d[i]) : Int64(0), 4)
function tensor!(p::Program, dims; dtype=F32, kind=T_TEMP, slot=-1, cval=0.0)
    push!(p.tensors, TensorDesc(dtype, length(dims), pad4(dims), kind, slot, cval))
    return length(p.tensors) - 1          # 0-based tensor id
end
const!(p, v) = tensor!(p, (); kind=T_CONST, cval=Float64(v))
dims(p::Program, t) = (d = p.tensors[t + 1]; d.dims[1:d.ndim])
function op!(p::Program, opcode, ins, outdims; iattr=(), fattr=())
    out = tensor!(p, outdims)
    ia = ntuple(i -> i <= length(iattr) ? Int64(iattr[i]) : Int64(0), 16)
    fa = ntuple(i -> i <= length(fattr) ? Float64(fattr[i]) : 0.0, 2)
    push!(p.ops, OpDesc(opcode, length(ins), ntuple(i -> i <= length(ins) ? Int32(ins[i]) : Int32(0), 6),
                        out, 0, ia, fa))
    return out
end

bcast_dims(a, b) = ntuple(i -> max(get(a, i, 1), get(b, i, 1)), max(length(a), length(b)))

"ChainRules `unbroadcast(x, Δ)` (spec: Yota src/helpers.jl:63-74)."
function unbroadcast!(p, xdims, d)
    dd = dims(p, d)
    xdims == dd && return d
    prod(xdims) == prod(dd) && return op!(p, OP.reshape, (d,), xdims)
    mask = 0
    for i in 1:length(dd)
        (i > length(xdims) || xdims[i] == 1) && (mask |= 1 << (i - 1))
    end
    return op!(p, OP.sum, (d,), xdims; iattr=(mask,), fattr=(1.0,))
end

# ---------------------------------------------------------------------------------------
# lowering of one rrule(cfg, f, args...) statement and of its pullback call
# (the rule math is ChainRules' / NNlib's; SURVEY.md §8(a) rows 7-12)
# ---------------------------------------------------------------------------------------
const UNARY = Dict{Any,Int}(tanh => EW.tanh, exp => EW.exp, log => EW.log, sin => EW.sin, cos => EW.cos,
                            abs2 => EW.square, sqrt => EW.sqrt, NNlib.relu => EW.relu, NNlib.σ => EW.sigmoid,
                            (-) => EW.neg)
const BINARY = Dict{Any,Int}((+) => EW.add, (-) => EW.sub, (*) => EW.mul, (/) => EW.div)

val_of(p, x) = x isa Number ? const!(p, x) : x

function lower_fwd!(p::Program, f, vals, kw, outdims)
    if f === broadcasted
        g = vals[1]
        if g === Base.literal_pow                         # x .^ 2
            x = vals[3]
            return op!(p, EW.square, (x,), outdims), (:sq, x)
        elseif g === (^)
            x, e = vals[2], Float64(vals[3])
            return op!(p, EW.powc, (x,), outdims; fattr=(e,)), (:powc, x, e)
        elseif length(vals) == 2 && haskey(UNARY, g)
            x = val_of(p, vals[2])
            y = op!(p, UNARY[g], (x,), outdims)
            return y, (:un, g, x, y)
        elseif haskey(BINARY, g)
            a, b = val_of(p, vals[2]), val_of(p, vals[3])
            y = op!(p, BINARY[g], (a, b), outdims)
            return y, (:bin, g, a, b, y)
        end
    elseif f === materialize
        return vals[1], (:id,)
    elseif f === (*)
        return op!(p, OP.matmul, (vals[1], vals[2]), outdims; iattr=(0, 0)), (:mm, vals[1], vals[2])
    elseif f === sum || f === Statistics.mean
        x = vals[1]; xd = dims(p, x)
        dd = haskey(kw, :dims) ? Tuple(kw[:dims]) : Tuple(1:length(xd))
        mask = sum(1 << (d - 1) for d in dd)
        n = prod(xd[d] for d in dd)
        y = op!(p, OP.sum, (x,), outdims; iattr=(mask,), fattr=(1.0,))
        f === Statistics.mean && (y = op!(p, EW.div, (y, const!(p, n)), outdims))
        return y, (:sum, xd, f === Statistics.mean ? n : nothing)
    elseif f === getindex
        x = vals[1]
        iattr, vecs = index_spec(vals[2:end])
        return op!(p, OP.getindex, (x, vecs...), outdims; iattr=iattr), (:idx, dims(p, x), iattr, vecs)
    elseif f === NNlib.logsoftmax
        get(kw, :dims, 1) == 1 || error("No derivative rule found for op logsoftmax(dims=$(kw[:dims])) on B200Array")
        y = op!(p, OP.logsoftmax, (vals[1],), outdims)
        return y, (:lsm, y)
    end
    # the analogue of src/grad.jl:179-180 -- there is no CPU fallback
    error("No derivative rule found for op $f on B200Array arguments (libyota_cuda has no lowering)")
end

function index_spec(I)
    iattr = Any[length(I)]; vecs = Int[]
    for ix in I
        if ix isa Colon
            append!(iattr, (IDX_COLON, 0, 0))
        elseif ix isa Integer
            append!(iattr, (IDX_INT, ix, ix))
        elseif ix isa AbstractUnitRange
            append!(iattr, (IDX_RANGE, first(ix), last(ix)))
        else                                              # tensor id of an Int64 vector
            append!(iattr, (IDX_VEC, length(vecs), 0)); push!(vecs, ix)
        end
    end
    return iattr, vecs
end

mul!(p, a, b) = op!(p, EW.mul, (a, b), bcast_dims(dims(p, a), dims(p, b)))

function lower_pb!(p::Program, f, saved, d)
    N = NoTangent()
    k = saved[1]
    if k === :id
        return (N, d)
    elseif k === :mm                                       # dA = Ȳ*B', dB = A'*Ȳ
        _, A, B = saved
        return (N, op!(p, OP.matmul, (d, B), dims(p, A); iattr=(0, 1)), op!(p, OP.matmul, (A, d), dims(p, B); iattr=(1, 0)))
    elseif k === :bin
        _, g, a, b, y = saved
        da, db = dims(p, a), dims(p, b)
        g === (+) && return (N, N, unbroadcast!(p, da, d), unbroadcast!(p, db, d))
        if g === (-)
            ub = unbroadcast!(p, db, d)
            return (N, N, unbroadcast!(p, da, d), op!(p, EW.neg, (ub,), dims(p, ub)))
        end
        g === (*) && return (N, N, unbroadcast!(p, da, mul!(p, d, b)), unbroadcast!(p, db, mul!(p, d, a)))
        ga = op!(p, EW.div, (d, b), dims(p, d))
        gb = op!(p, EW.neg, (mul!(p, d, op!(p, EW.div, (y, b), dims(p, y))),), dims(p, d))
        return (N, N, unbroadcast!(p, da, ga), unbroadcast!(p, db, gb))
    elseif k === :sq                                       # 2 .* Δ .* x
        x = saved[2]
        return (N, N, N, mul!(p, op!(p, EW.mulc, (d,), dims(p, d); fattr=(2.0,)), x), N)
    elseif k === :powc
        _, x, e = saved; xd = dims(p, x)
        g = op!(p, EW.mulc, (op!(p, EW.powc, (x,), xd; fattr=(e - 1,)),), xd; fattr=(e,))
        return (N, N, mul!(p, d, g), N)
    elseif k === :un
        _, g, x, y = saved; xd = dims(p, x)
        dg = g === tanh ? op!(p, EW.rsubc, (op!(p, EW.square, (y,), xd),), xd; fattr=(1.0,)) :
             g === exp ? y :
             g === log ? op!(p, EW.rdivc, (x,), xd; fattr=(1.0,)) :
             g === sin ? op!(p, EW.cos, (x,), xd) :
             g === cos ? op!(p, EW.neg, (op!(p, EW.sin, (x,), xd),), xd) :
             g === abs2 ? op!(p, EW.mulc, (x,), xd; fattr=(2.0,)) :
             g === sqrt ? op!(p, EW.rdivc, (op!(p, EW.mulc, (y,), xd; fattr=(2.0,)),), xd; fattr=(1.0,)) :
             g === NNlib.relu ? op!(p, EW.gt0, (x,), xd) :
             g === NNlib.σ ? mul!(p, y, op!(p, EW.rsubc, (y,), xd; fattr=(1.0,))) : nothing
        t = g === (-) ? op!(p, EW.neg, (d,), dims(p, d)) : mul!(p, d, dg)
        return (N, N, unbroadcast!(p, xd, t))
    elseif k === :sum                                      # broadcast-back (÷ n for mean)
        _, xd, n = saved
        n !== nothing && (d = op!(p, EW.div, (d, const!(p, n)), dims(p, d)))
        return (N, op!(p, EW.copy, (d,), xd))
    elseif k === :idx
        _, xd, iattr, vecs = saved
        return (N, op!(p, OP.ungetindex, (d, vecs...), xd; iattr=iattr), ntuple(_ -> N, iattr[1])...)
    elseif k === :lsm                                      # dy .- sum(dy; dims=1) .* exp.(y)
        y = saved[2]; yd = dims(p, y)
        s = op!(p, OP.sum, (d,), (1, yd[2:end]...); iattr=(1,), fattr=(1.0,))
        return (N, op!(p, EW.sub, (d, mul!(p, s, op!(p, EW.exp, (y,), yd))), yd))
    end
    error("no pullback lowering for $f")
end

# ---------------------------------------------------------------------------------------
# Tape{GradCtx} -> program.  Statement forms: SURVEY.md §3.2.
# ---------------------------------------------------------------------------------------
struct Compiled
    h::Ptr{Cvoid}
    arg_slots::Vector{Int}            # per tape input: in slot (or 0 for non-array arguments)
    out_dims::Vector{Tuple}
    val::Any                          # out slot of the value
    grads::Vector{Any}                # per input: out slot | ZeroTangent() | NoTangent()
    default_seed::Float32
end

function lower(tape::Tape{GradCtx})
    p = Program()
    env = Dict{Int,Any}()
    arg_slots = Int[]
    seed_id = tape.meta[:seed].id
    for op in tape.ops
        i = op.id
        if op isa Input
            if op.val isa B200Array
                env[i] = tensor!(p, size(op.val); dtype=eltype(op.val) == Int64 ? I64 : F32, kind=T_INPUT, slot=p.n_in)
                push!(arg_slots, p.n_in); p.n_in += 1
            else
                env[i] = op.val; push!(arg_slots, -1)     # the function object, numbers
            end
        elseif op isa Constant
            env[i] = i == seed_id ? tensor!(p, (); kind=T_SEED) : op.val
        elseif op isa Call
            fn = op.fn isa V ? tape[op.fn].val : op.fn
            if fn === rrule                               # rrule(cfg, f, args...)
                f = op.args[2] isa V ? tape[op.args[2]].val : op.args[2]
                vals = [a isa V ? env[a.id] : a for a in op.args[3:end]]
                outdims = size(op.val[1])
                env[i] = (:rr, f, lower_fwd!(p, f, vals, NamedTuple(), outdims)...)
            elseif fn === Core.kwfunc(rrule)              # kw call: (kw, rrule, cfg, f, args...)  src/grad.jl:143-145
                kw = op.args[1] isa V ? tape[op.args[1]].val : op.args[1]
                f = op.args[4] isa V ? tape[op.args[4]].val : op.args[4]
                vals = [a isa V ? env[a.id] : a for a in op.args[5:end]]
                env[i] = (:rr, f, lower_fwd!(p, f, vals, kw, size(op.val[1]))...)
            elseif fn === _getfield
                rr = env[op.args[1].id]
                env[i] = op.args[2] == 1 ? rr[3] : (:pb, rr[2], rr[4])
            elseif op.fn isa V && env[op.fn.id] isa Tuple && env[op.fn.id][1] === :pb      # pb(dy)
                _, f, saved = env[op.fn.id]
                env[i] = lower_pb!(p, f, saved, val_of(p, env[op.args[1].id]))
            elseif fn === getfield
                env[i] = env[op.args[1].id][op.args[2]]
            elseif fn === broadcast && op.args[1] === (+)  # tape gradient accumulation, src/grad.jl:91
                a, b = env[op.args[2].id], env[op.args[3].id]
                env[i] = a isa NoTangent || a isa ZeroTangent ? b : b isa NoTangent || b isa ZeroTangent ? a :
                         op!(p, EW.add, (a, b), bcast_dims(dims(p, a), dims(p, b)))
            elseif fn === tuple
                env[i] = Tuple(a isa V ? env[a.id] : a for a in op.args)
            elseif fn === map                              # map(unthunk, t): thunks are plain dataflow here
                env[i] = env[op.args[2].id]
            else
                error("No derivative rule found for op $op on B200Array arguments")
            end
        end
    end
    val, grads = env[tape.result.id]
    out_dims = Tuple[]
    function out!(x)
        (x isa ZeroTangent || x isa NoTangent) && return x
        d = p.tensors[x + 1]
        if d.kind != T_TEMP                                # inputs/seed: copy so every output has one writer
            x = op!(p, EW.copy, (x,), d.dims[1:d.ndim]); d = p.tensors[x + 1]
        end
        p.tensors[x + 1] = TensorDesc(d.dtype, d.ndim, d.dims, T_OUTPUT, p.n_out, 0.0)
        push!(out_dims, d.dims[1:d.ndim]); p.n_out += 1
        return p.n_out - 1
    end
    v = out!(val)
    gs = Any[out!(g) for g in grads]
    return p, arg_slots, out_dims, v, gs
end

"Replacement for `Yota.grad_compile(tape)` (src/grad.jl:293-307) when the arguments are B200Arrays."
function b200_compile(tape::Tape{GradCtx}; flags::UInt32=FLAG_GEMM_TF32)
    p, arg_slots, out_dims, v, gs = lower(tape)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:yota_program_build, LIB), Cint,
                (Ptr{Cvoid}, Ptr{OpDesc}, Int64, Ptr{TensorDesc}, Int64, UInt32, Ptr{Ptr{Cvoid}}),
                context().h, p.ops, length(p.ops), p.tensors, length(p.tensors), flags, h))
    seed = tape[tape.meta[:seed]].val
    c = Compiled(h[], arg_slots, out_dims, v, gs, seed isa Number ? Float32(seed) : 1f0)
    # same calling convention as the function grad_compile returns: gf(f, args...; seed)
    return function (f, args...; seed=c.default_seed)
        ins = Vector{Ptr{Cvoid}}(undef, count(>=(0), c.arg_slots))
        for (a, s) in zip((f, args...), c.arg_slots)
            s >= 0 && (ins[s + 1] = a.ptr)
        end
        outs = [B200Array{Float32,length(d)}(Tuple(Int.(d))) for d in c.out_dims]
        optrs = Ptr{Cvoid}[o.ptr for o in outs]
        check(ccall((:yota_program_run, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Ptr{Cvoid}}, Int64, Cfloat),
                    c.h, ins, length(ins), optrs, length(optrs), Float32(seed)))
        pick(x) = x isa Integer ? outs[x + 1] : x
        val = pick(c.val)
        ndims(val) == 0 && (val = Array(val)[])            # scalar loss: one 4-byte read back
        return val, map(pick, Tuple(c.grads))
    end
end

# ---------------------------------------------------------------------------------------
# Operator seam (SURVEY.md §8(b), Seam 2): the primitives and their rrules on B200Arrays.
#
# Umlaut's tracer and Yota's tape builder EXECUTE every statement they record (`mkcall`:
# the primal call while tracing, `rrule(cfg, f, args...)` in chainrules_transform!,
# src/grad.jl:145-147, and every `pb(dy)` in back!, :173), which is how a cold `grad` call
# returns `tape.retval` (:357-360).  With B200Array arguments those calls dispatch to the methods
# below: each one builds the single-rule op program with the same `lower_fwd!` / `lower_pb!` the
# fused path uses and runs it on the device (yota_program_build + yota_program_run, one launch
# list per statement, no graph) -- the cold path never touches host copies of the data.
# Python mirror (tested on the device): yota.jl_b200/eager.py `rrule` / `play`.
# ---------------------------------------------------------------------------------------
const EAGER_CACHE = Dict{Any,Any}()

"Build (once per signature) and run a program whose inputs are `arrays`; returns the outputs."
function run_program(key, build::Function, arrays::Vector)
    ent = get!(EAGER_CACHE, key) do
        p, out_dims, extra = build()
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:yota_program_build, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{OpDesc}, Int64, Ptr{TensorDesc}, Int64, UInt32, Ptr{Ptr{Cvoid}}),
                    context().h, p.ops, length(p.ops), p.tensors, length(p.tensors), FLAG_NO_GRAPH, h))
        (h[], out_dims, extra)
    end
    h, out_dims, extra = ent
    ins = Ptr{Cvoid}[a.ptr for a in arrays]
    outs = [B200Array{Float32,length(d)}(Tuple(Int.(d))) for d in out_dims]
    optrs = Ptr{Cvoid}[o.ptr for o in outs]
    check(ccall((:yota_program_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Ptr{Cvoid}}, Int64, Cfloat),
                h, ins, length(ins), optrs, length(optrs), 1f0))
    return outs, extra
end

sig(a::B200Array) = (eltype(a), size(a))
sig(a) = a

"Result dims of `f(args...; kw...)` on arrays of these sizes (what Julia itself would return)."
function out_dims(f, args, kw)
    sz(a) = a isa B200Array ? size(a) : ()
    f === broadcasted && return Base.Broadcast.broadcast_shape(map(sz, args[2:end])...)
    f === materialize && return sz(args[1])
    f === (*) && return (size(args[1], 1), size(args[2])[2:end]...)
    if f === sum || f === Statistics.mean
        haskey(kw, :dims) || return ()
        return ntuple(i -> i in kw[:dims] ? 1 : size(args[1], i), ndims(args[1]))
    end
    f === getindex && return map(length, Base.index_shape(Base.to_indices(args[1], map(i -> i isa B200Array ? (1:length(i)) : i, args[2:end]))...))
    f === NNlib.logsoftmax && return sz(args[1])
    error("No derivative rule found for op $f on B200Array arguments (libyota_cuda has no lowering)")
end

"""
    eager_rrule(f, args...; kw...) -> (value, pullback)

`rrule(YotaRuleConfig(), f, args...)` for device arrays: the forward runs now as one tiny device
program; `pullback(dy)` runs the rule's pullback as another and returns the tangent tuple.
"""
function eager_rrule(f, args...; kw...)
    arrays = B200Array[a for a in args if a isa B200Array]
    od = out_dims(f, args, kw)
    outs, (pfwd, y, saved) = run_program((:fwd, f, map(sig, args), kw), arrays) do
        p = Program()
        vals = Any[a isa B200Array ? (t = tensor!(p, size(a); dtype=eltype(a) == Int64 ? I64 : F32, kind=T_INPUT, slot=p.n_in); p.n_in += 1; t) : a
                   for a in args]
        y, saved = lower_fwd!(p, f, vals, kw, od)
        mark_output!(p, y)
        (p, [od], (p, y, saved))
    end
    val = outs[1]
    return val, dy -> run_pullback(f, args, kw, pfwd, y, saved, arrays, val, dy)
end

function mark_output!(p::Program, x)
    d = p.tensors[x + 1]
    if d.kind != T_TEMP
        x = op!(p, EW.copy, (x,), d.dims[1:d.ndim]); d = p.tensors[x + 1]
    end
    p.tensors[x + 1] = TensorDesc(d.dtype, d.ndim, d.dims, T_OUTPUT, p.n_out, 0.0)
    p.n_out += 1
    return x
end

"Tangents of one rule on the device: builds/caches the pullback program, binds dy / the value / the arguments."
function pullback_program(f, args, kw, pfwd::Program, y, saved, dyv)
    p = Program()
    binds = Any[]
    memo = Dict{Int,Int}()
    inp(dims, dtype, what) = (t = tensor!(p, dims; dtype=dtype, kind=T_INPUT, slot=p.n_in); p.n_in += 1; push!(binds, what); t)
    function remap(x)
        if x isa Integer && 0 <= x < length(pfwd.tensors)   # a tensor id of the forward program
            haskey(memo, x) && return memo[x]
            d = pfwd.tensors[x + 1]
            r = d.kind == T_CONST ? const!(p, d.cval) :
                d.kind == T_INPUT ? inp(d.dims[1:d.ndim], d.dtype, d.slot + 1) : inp(d.dims[1:d.ndim], d.dtype, :val)
            return memo[x] = r
        end
        return x isa Tuple ? map(remap, x) : x isa Vector ? map(remap, x) : x
    end
    # which entries of `saved` are tensor ids depends on the rule (dims tuples and counts are Ints too)
    k = saved[1]
    saved2 = k === :sum ? saved :
             k === :idx ? (k, saved[2], saved[3], map(remap, saved[4])) :
             k === :powc ? (k, remap(saved[2]), saved[3]) :
             (k, map(x -> x isa Integer ? remap(x) : x, saved[2:end])...)
    d2 = dyv isa B200Array ? inp(size(dyv), F32, :dy) : const!(p, dyv)
    tangents = lower_pb!(p, f, saved2, d2)
    out_dims_ = Tuple[]
    for t in tangents
        t isa Integer || continue
        mark_output!(p, t); push!(out_dims_, dims(p, t))
    end
    return p, out_dims_, (binds, tangents)
end

function run_pullback(f, args, kw, pfwd::Program, y, saved, arrays, val, dy)
    (dy isa ZeroTangent || dy isa NoTangent) && return ntuple(_ -> ZeroTangent(), length(args) + 1)
    dyv = unthunk(dy)
    key = (:pb, f, map(sig, args), kw, sig(dyv))
    ent = get(EAGER_CACHE, key, nothing)
    binds = ent === nothing ? pullback_program(f, args, kw, pfwd, y, saved, dyv)[3][1] : ent[3][1]
    ins = B200Array[w === :dy ? dyv : w === :val ? val : arrays[w] for w in binds]
    outs, (_, tangents) = run_program(() -> pullback_program(f, args, kw, pfwd, y, saved, dyv), key, ins)
    k = 0
    return map(t -> t isa Integer ? outs[k += 1] : t, tangents)
end
run_program(build::Function, key, arrays) = run_program(key, build, arrays)

# the primitives themselves (what `mkcall` executes while TRACING) and their rules
const DeviceRules = Union{typeof(*),typeof(sum),typeof(Statistics.mean),typeof(getindex),typeof(NNlib.logsoftmax)}
for f in (:(Base.:*), :(Base.sum), :(Statistics.mean), :(NNlib.logsoftmax))
    @eval $f(a::B200Array, rest...; kw...) = eager_rrule($f, a, rest...; kw...)[1]
    @eval ChainRulesCore.rrule(::YotaRuleConfig, ::typeof($f), a::B200Array, rest...; kw...) =
        eager_rrule($f, a, rest...; kw...)
end
Base.getindex(a::B200Array, I::Union{Colon,Integer,AbstractUnitRange,B200Array{Int64,1}}...) = eager_rrule(getindex, a, I...)[1]
ChainRulesCore.rrule(::YotaRuleConfig, ::typeof(getindex), a::B200Array, I...) = eager_rrule(getindex, a, I...)
# dotted calls: a broadcast over device arrays is executed at once (the "lazy" Broadcasted of the
# reference's fast rules is an array here; `materialize` is the identity both ways, src/rulesets.jl:27-29)
Base.Broadcast.broadcasted(f, a::B200Array, rest...) = eager_rrule(broadcasted, f, a, rest...)[1]
Base.Broadcast.broadcasted(f, x::Number, a::B200Array, rest...) = eager_rrule(broadcasted, f, x, a, rest...)[1]
Base.Broadcast.materialize(a::B200Array) = a
ChainRulesCore.rrule(::YotaRuleConfig, ::typeof(broadcasted), f, a::B200Array, rest...) = eager_rrule(broadcasted, f, a, rest...)
ChainRulesCore.rrule(::YotaRuleConfig, ::typeof(broadcasted), f, x::Number, a::B200Array, rest...) =
    eager_rrule(broadcasted, f, x, a, rest...)
ChainRulesCore.rrule(::YotaRuleConfig, ::typeof(materialize), a::B200Array) = (a, dy -> (NoTangent(), dy))
Base.broadcast(::typeof(+), a::B200Array, b::B200Array) = eager_rrule(broadcasted, +, a, b)[1]   # src/grad.jl:91

# Hook: `Yota.grad` on device arrays.  The cold call traces and builds the tape ON THE DEVICE
# through the methods above (no host copies); the finished tape is compiled into the fused
# program, which serves every warm call (GRAD_CACHE hit, src/grad.jl:353-355).
function Yota.grad(f, args::Vararg{Union{B200Array,Number}}; seed=1)
    key = (f, map(a -> a isa B200Array ? (eltype(a), size(a)) : typeof(a), args))  # sizes too: the program is shape-specialised
    if haskey(Yota.GRAD_CACHE, key)
        return Base.invokelatest(Yota.GRAD_CACHE[key], f, args...; seed=seed)
    end
    tape = Yota.gradtape(f, args...; seed=seed)          # executes statement by statement on the device
    Yota.GRAD_CACHE[key] = b200_compile(tape)
    return tape.retval                                   # src/grad.jl:357-360
end

"update!(x, gx): x .= x .- gx on the device (src/update.jl:42-49)."
function Yota.update!(x::B200Array{Float32}, gx::B200Array{Float32}, lr::Real=1)
    check(ccall((:yota_update_sgd, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cfloat),
                context().h, x.ptr, gx.ptr, length(x), Float32(lr)))
    return x
end

# ---- multi-GPU: one Julia process per GPU (INTEGRATION.md) ---------------------------------
"Create the communicator of this process' context from a 128-byte id made on rank 0 (`comm_unique_id`)."
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:yota_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init(rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    check(ccall((:yota_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), context().h, rank, nranks, id))

"""
    symmetric_array(T, dims...)

Collective: an output array in symmetric memory (same call, same sizes, same order on every
rank).  Arrays allocated this way and marked with `set_allreduce!` are reduced by the library's
own NVSwitch kernel instead of `ncclAllReduce`.  Falls back to an ordinary `B200Array` when
symmetric memory is unavailable.
"""
function symmetric_array(::Type{Float32}, dims::Integer...)
    ctx = context()
    if ccall((:yota_comm_symm_available, LIB), Cint, (Ptr{Cvoid},), ctx.h) == 0
        return B200Array{Float32,length(dims)}(Tuple(Int.(dims)))
    end
    p = Ref{Ptr{Cvoid}}()
    check(ccall((:yota_comm_symm_alloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}), ctx.h, 4 * prod(dims), p))
    a = B200Array{Float32,length(dims)}(p[], Tuple(Int.(dims)))   # non-owning constructor: freed with yota_comm_symm_free
    finalizer(x -> ccall((:yota_comm_symm_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, x.ptr), a)
    return a
end

"Mark program outputs (0-based out slots) whose buffers are summed over ranks inside every run."
set_allreduce!(prog::Ptr{Cvoid}, slots::Vector{Int32}) =
    check(ccall((:yota_program_set_allreduce, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int64), prog, slots, length(slots)))

end # module
