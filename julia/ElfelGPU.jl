# ElfelGPU.jl -- Julia-side shim selecting the B200 assembler in place of SysmatAssemblerSparse.
#
# SHIPPED UNEXECUTED: no Julia runtime exists in the build/test image, so this file has never been run.
# It binds exactly the C ABI of include/elfel_gpu.h, which IS exercised (through Python ctypes) by tests/
# and bench.py.  Drop it next to a checkout of Elfel.jl and `include("ElfelGPU.jl")`.
module ElfelGPU

using SparseArrays: SparseMatrixCSC
using Elfel.Assemblers: AbstractSysmatAssembler, AbstractSysvecAssembler
import Elfel.Assemblers: start!, assemble!, finish!
using Elfel.FEIterators: FEIterator
using Elfel.QPIterators: QPIterator
using Elfel.RefShapes: npts

const LIB = get(ENV, "ELFELGPU_LIB", "libelfelgpu.so")

# weak forms = the integrate! closures of the examples
struct HeatForm;            kappa::Float64; end                 # examples/heat/poisson/t3.jl:53-58
struct HeatLoadForm;        Q::Float64; end                     # examples/heat/poisson/t3.jl:57 (vector form)
struct ElasticityForm;      D::Matrix{Float64}; end             # examples/elasticity/stretch/t6.jl:42-58
struct StokesGenForm;       D::Matrix{Float64}; end             # examples/stokes/colliding_flow/ht_p2_p1_gen.jl
struct StokesReddyForm;     mu::Float64; end                    # .../ht_p2_p1.jl
struct StokesVeclapAltForm; mu::Float64; end                    # .../ht_p2_p1_veclap_alt.jl
struct StokesVeclapForm;    mu::Float64; end                    # .../ht_p2_p1_veclap.jl
formid(::HeatForm) = 1; formid(::ElasticityForm) = 2; formid(::StokesGenForm) = 3
formid(::StokesReddyForm) = 4; formid(::StokesVeclapAltForm) = 5; formid(::StokesVeclapForm) = 6
params(f::HeatForm) = [f.kappa]
params(f::Union{ElasticityForm,StokesGenForm}) = vec(collect(Float64, f.D))     # column-major 3x3
params(f::Union{StokesReddyForm,StokesVeclapAltForm,StokesVeclapForm}) = [f.mu]

mutable struct SysmatAssemblerGPU <: AbstractSysmatAssembler
    ctx::Ptr{Cvoid}
    nrow::Int64
    ncol::Int64
    nnz::Int64
    colptr::Vector{Int64}      # the SparseMatrixCSC fields: allocated in assemble! as soon as nnz is known, so that the
    rowval::Vector{Int64}      # structure is already on its way to the host while the values are computed
    nzval::Vector{Float64}
    loaded::Vector{Any}        # the FEIterators of the last assemble! call, in space-slot order
    function SysmatAssemblerGPU(zero::Float64 = 0.0; device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:efg_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
        rc == 0 || error("efg_create failed ($rc): no usable CUDA device (there is no CPU fallback)")
        a = new(r[], 0, 0, 0, Int64[], Int64[], Float64[], Any[])
        # EFG_OPT_DEFER_XY (6): the coordinate array is borrowed until efg_pattern returns and copied while the pattern kernels
        # run; assemble! below keeps the iterators (and with them the mesh) alive across the whole sequence
        ccall((:efg_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), a.ctx, 6, 1)
        finalizer(x -> ccall((:efg_destroy, LIB), Cint, (Ptr{Cvoid},), x.ctx), a)
        return a
    end
end

function _check(a::SysmatAssemblerGPU, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:efg_last_error, LIB), Cstring, (Ptr{Cvoid},), a.ctx))
    rc == -4 ? throw(ArgumentError(msg)) : error("libelfelgpu error $rc: $msg")
end

"start!(ass, nrow, ncol) -- src/Assemblers.jl:67-78"
function start!(a::SysmatAssemblerGPU, nrow, ncol)
    a.nrow, a.ncol = nrow, ncol
    _check(a, ccall((:efg_start, LIB), Cint, (Ptr{Cvoid}, Int64, Int64), a.ctx, nrow, ncol))
    return a
end

_kind(it::FEIterator) = length(it._nodes)            # 3 = T3, 4 = Q4, 6 = T6
_rule(q::QPIterator, kind) = kind == 4 ? isqrt(npts(q._quadr)) : npts(q._quadr)

"""
    assemble!(ass, form, elits, qpits)

Replaces the whole `for el in elit ... assemble!(ass, ke) end` loop of the examples.  `elits`/`qpits` are
one iterator (heat, elasticity) or a tuple in space order (Stokes: (uel, pel) or (uxel, uyel, pel)).
"""
function assemble!(a::SysmatAssemblerGPU, form, elits, qpits)
    elits isa FEIterator && (elits = (elits,); qpits = (qpits,))
    meshes = Any[]
    for it in elits
        any(m -> m === it._bir, meshes) || push!(meshes, it._bir)
    end
    GC.@preserve elits begin
        for (slot, it) in enumerate(elits)
            mslot = findfirst(m -> m === it._bir, meshes)
            if count(j -> elits[j]._bir === it._bir, 1:slot) == 1      # first space on this mesh: send the mesh
                conn = reinterpret(Int64, it._bir._v)                  # nen x nel, contiguous SVector storage
                xy = reinterpret(Float64, it._geom.v)                  # 2 x nnodes
                _check(a, ccall((:efg_set_mesh, LIB), Cint,
                                (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Int64}, Ptr{Float64}),
                                a.ctx, mslot - 1, _kind(it), length(it), length(it._geom), conn, xy))
            end
            if it._fld2 === nothing                                     # vertex dofs only: FEH1_T3 / FEH1_Q4 / FEH1_T6
                dof = reinterpret(Int64, it._fld0.dofnums)             # ncomp x nnodes
                ncomp = length(eltype(it._fld0.dofnums))
                _check(a, ccall((:efg_set_space, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Ptr{Int64}),
                                a.ctx, slot - 1, mslot - 1, ncomp, length(it._fld0.dofnums), dof))
            else                                                       # a dof on the cell: FEH1_T3_BUBBLE (7), FEL2_T3 / FEL2_Q4 (1)
                fe = it._fld0 === nothing ? 1 : 7                      # EFG_FE_L2 / EFG_FE_T3_BUBBLE (src/FEIterators.jl:66-79)
                cdof = reinterpret(Int64, it._fld2.dofnums)            # ncomp x nel
                ncomp = length(eltype(it._fld2.dofnums))
                ndof = it._fld0 === nothing ? C_NULL : pointer(reinterpret(Int64, it._fld0.dofnums))
                nn = it._fld0 === nothing ? 0 : length(it._fld0.dofnums)
                _check(a, ccall((:efg_set_space_fe, LIB), Cint,
                                (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Int64, Ptr{Int64}, Int64, Ptr{Int64}),
                                a.ctx, slot - 1, mslot - 1, fe, ncomp, nn, ndof, length(it._fld2.dofnums), cdof))
            end
        end
        _check(a, ccall((:efg_start, LIB), Cint, (Ptr{Cvoid}, Int64, Int64), a.ctx, a.nrow, a.ncol))
        p = params(form)
        nnz = Ref{Int64}(0)
        rule = _rule(qpits[1], _kind(elits[1]))
        # structure first: the pattern (-> nnz) ...
        _check(a, ccall((:efg_pattern, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Int64}), a.ctx, formid(form), rule, nnz))
        a.nnz = nnz[]
        a.colptr = Vector{Int64}(undef, a.ncol + 1)
        a.rowval = Vector{Int64}(undef, a.nnz)
        a.nzval = Vector{Float64}(undef, a.nnz)
        # ... starts travelling to the host (copy stream) while the scatter maps and the values are computed
        _check(a, ccall((:efg_fetch_pattern_async, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), a.ctx, a.colptr, a.rowval))
        _check(a, ccall((:efg_numeric, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), a.ctx, p, length(p)))
        a.loaded = Any[it for it in elits]
    end
    return a
end

"finish!(ass) -- src/Assemblers.jl:121-123: waits for the structure, copies the values; the arrays are wrapped without a copy."
function finish!(a::SysmatAssemblerGPU)
    _check(a, ccall((:efg_fetch_csc, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                    a.ctx, C_NULL, C_NULL, a.nzval))
    return SparseMatrixCSC(a.nrow, a.ncol, a.colptr, a.rowval, a.nzval)
end

# ---- system vector (SysvecAssembler, src/Assemblers.jl:170-232) ---------------------------------------------
"""
    SysvecAssemblerGPU(0.0; like = am)

Selected in place of `SysvecAssembler`.  `like = am` shares the device context of a `SysmatAssemblerGPU`, so the
mesh and dof maps sent for the matrix are reused by the vector half of the same `integrate!` call.
"""
mutable struct SysvecAssemblerGPU <: AbstractSysvecAssembler
    owner::SysmatAssemblerGPU
    ndofs::Int64
    SysvecAssemblerGPU(zero::Float64 = 0.0; like::SysmatAssemblerGPU = SysmatAssemblerGPU(0.0)) = new(like, 0)
end

"start!(av, nrow) -- src/Assemblers.jl:196-200"
start!(v::SysvecAssemblerGPU, nrow::Int64) = (v.ndofs = nrow; v)

"assemble!(av, HeatLoadForm(Q), elit, qpit): the `init!(fe, eldofs(el)); fe[j] += N[j]*Q*JxW; assemble!(av, fe)` half of the loop"
function assemble!(v::SysvecAssemblerGPU, form::HeatLoadForm, elit::FEIterator, qpit::QPIterator)
    a = v.owner       # mesh 0 / space 0 were sent by assemble!(am, HeatForm(...), elit, qpit) on the shared context
    p = [form.Q]
    _check(a, ccall((:efg_vec_assemble, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Cint, Int64),
                    a.ctx, 1, _rule(qpit, _kind(elit)), p, 1, v.ndofs))
    return v
end

"""
    assemble!(am, av, HeatForm(kappa), HeatLoadForm(Q), elit, qpit)

K and F of ONE pass, like the reference's heat loops (`ke[i, j] += ...` and `fe[j] += N[j] * Q * JxW` in the same quadrature
loop, `assemble!(am, ke); assemble!(av, fe)` per element: examples/heat/poisson/t3.jl:41-64).  `av` must share `am`'s
context (`SysvecAssemblerGPU(0.0; like = am)`); then `finish!(am)`, `finish!(av)` as usual.
"""
function assemble!(a::SysmatAssemblerGPU, v::SysvecAssemblerGPU, form::HeatForm, vform::HeatLoadForm, elit::FEIterator, qpit::QPIterator)
    v.owner === a || error("assemble!(am, av, ...): the vector assembler must share the matrix assembler's context (like = am)")
    _check(a, ccall((:efg_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), a.ctx, 5, 1))      # EFG_OPT_FUSE_LOAD (before the symbolic phase)
    GC.@preserve elit begin
        conn = reinterpret(Int64, elit._bir._v); xy = reinterpret(Float64, elit._geom.v)
        _check(a, ccall((:efg_set_mesh, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Int64}, Ptr{Float64}),
                        a.ctx, 0, _kind(elit), length(elit), length(elit._geom), conn, xy))
        dof = reinterpret(Int64, elit._fld0.dofnums)
        _check(a, ccall((:efg_set_space, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Ptr{Int64}),
                        a.ctx, 0, 0, 1, length(elit._fld0.dofnums), dof))
        _check(a, ccall((:efg_start, LIB), Cint, (Ptr{Cvoid}, Int64, Int64), a.ctx, a.nrow, a.ncol))
        nnz = Ref{Int64}(0)
        _check(a, ccall((:efg_pattern, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Int64}), a.ctx, 1, _rule(qpit, _kind(elit)), nnz))
        a.nnz = nnz[]
        a.colptr = Vector{Int64}(undef, a.ncol + 1); a.rowval = Vector{Int64}(undef, a.nnz); a.nzval = Vector{Float64}(undef, a.nnz)
        _check(a, ccall((:efg_fetch_pattern_async, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), a.ctx, a.colptr, a.rowval))
        _check(a, ccall((:efg_numeric_with_load, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Float64), a.ctx, [form.kappa], 1, vform.Q))
        a.loaded = Any[elit]
    end
    return a, v
end

"finish!(av) -- src/Assemblers.jl:230-232"
function finish!(v::SysvecAssemblerGPU)
    F = Vector{Float64}(undef, v.ndofs)
    _check(v.owner, ccall((:efg_fetch_vec, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), v.owner.ctx, F))
    return F
end

# ---- what solve! does with K right after finish! (examples/heat/poisson/t3.jl:77-80), without fetching K ------
"`KT = mul(am, T)` == `K * T`, same summation order as SparseArrays"
function mul(a::SysmatAssemblerGPU, x::Vector{Float64})
    y = Vector{Float64}(undef, a.nrow)
    _check(a, ccall((:efg_spmv, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), a.ctx, x, y))
    return y
end

"`block(am, 1:nu, 1:nu)` == `K[1:nu, 1:nu]`, sliced on the device"
function block(a::SysmatAssemblerGPU, rows::UnitRange{Int}, cols::UnitRange{Int})
    n = Ref{Int64}(0)
    _check(a, ccall((:efg_block_nnz, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ref{Int64}),
                    a.ctx, first(rows), last(rows), first(cols), last(cols), n))
    colptr = Vector{Int64}(undef, length(cols) + 1)
    rowval = Vector{Int64}(undef, n[])
    nzval = Vector{Float64}(undef, n[])
    _check(a, ccall((:efg_fetch_block, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                    a.ctx, colptr, rowval, nzval))
    return SparseMatrixCSC(length(rows), length(cols), colptr, rowval, nzval)
end

# ---- post-processing integrators (examples/stokes/colliding_flow/ht_p2_p1.jl:120-178) -------------------------
"""
    evaluate_error(am, elits, qpit, U, truefs)

`sqrt(sum_el sum_qp JxW * sum_c (u_c(qp) - truef_c(location(el, qp)...))^2)`: evaluate_pressure_error /
evaluate_velocity_error with the element loop on the device.  `elits`: one FEIterator per field component (the
same iterator twice for the two components of a vector space), all of them among the iterators of the last
`assemble!(am, ...)` call; `truefs`: one function per component.  Same signature as the Python mirror
(elfel.jl_b200/assemblers.py:evaluate_error) and as INTEGRATION.md shows.
"""
function evaluate_error(a::SysmatAssemblerGPU, elits, qpit::QPIterator, U::Vector{Float64}, truefs)
    elits isa FEIterator && (elits = (elits,); truefs = (truefs,))
    N = length(elits)
    slots = Tuple{Int,Int}[]
    seen = Dict{Int,Int}()
    for it in elits                                   # (space slot, component) of every field component
        s = findfirst(k -> k._fld0 === it._fld0, a.loaded)
        s === nothing && error("evaluate_error: this space was not part of the last assemble! call")
        c = get(seen, s, 0)
        push!(slots, (s - 1, c))
        seen[s] = c + 1
    end
    mesh_slot = a.loaded[1]._bir === elits[1]._bir ? 0 : 1
    nel = length(elits[1])
    quad = _rule(qpit, _kind(elits[1]))
    np = Ref{Int64}(0)
    _check(a, ccall((:efg_qp_locations, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ref{Int64}), a.ctx, mesh_slot, quad, C_NULL, np))
    loc = Array{Float64}(undef, 2, np[], nel)
    _check(a, ccall((:efg_qp_locations, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ref{Int64}), a.ctx, mesh_slot, quad, loc, np))
    truth = Array{Float64}(undef, N, np[], nel)
    for c in 1:N
        truth[c, :, :] .= truefs[c].(view(loc, 1, :, :), view(loc, 2, :, :))
    end
    ss = Cint[s[1] for s in slots]; cc = Cint[s[2] for s in slots]
    out = Ref{Float64}(0.0)
    _check(a, ccall((:efg_l2_error, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Ref{Float64}),
                    a.ctx, N, ss, cc, quad, U, length(U), truth, out))
    return out[]
end

end # module
