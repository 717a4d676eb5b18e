"""
    SpaceChargeB200

Drop-in replacement for the hot path of SpaceCharge.jl v1.2.0 on NVIDIA B200 (sm_100a).

Same exported surface as the reference (`src/SpaceCharge.jl:17`): `Mesh3D, deposit!, clear_mesh!,
interpolate_field, solve!`, same positional/keyword arguments, same `ErrorException`s.  Every
computation is a `ccall` into `libspacecharge_b200.so` (C ABI: `include/spacecharge_b200.h`);
there is no CPU backend and no KernelAbstractions dispatch.  Arrays are `CuArray`s owned by Julia.

NOTE: Julia is not installed in the build image, so this shim has never been executed there; it is
kept mechanical (argument marshalling only).  The tested twin of this file is the Python/ctypes
mirror `spacecharge.jl_b200/__init__.py`, which binds exactly the same symbols.
"""
module SpaceChargeB200

using CUDA

export Mesh3D, deposit!, clear_mesh!, interpolate_field, solve!

const CLIGHT = 299792458.0               # src/utils.jl:7
const FPEI = CLIGHT^2 * 1.0e-7           # src/utils.jl:8

const LIB = get(ENV, "SPACECHARGE_B200_LIB", "libspacecharge_b200.so")

const SCB_F32 = Cint(0)
const SCB_F64 = Cint(1)
dtag(::Type{Float32}) = SCB_F32
dtag(::Type{Float64}) = SCB_F64

# ---- handle (one per task/stream, like the C ABI requires) --------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    stream::Ptr{Cvoid}     # the CUDA stream the handle currently enqueues on
end

function Handle(; device::Integer = CUDA.deviceid(CUDA.device()), stream = CUDA.stream())
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:scb_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
               device, Base.unsafe_convert(Ptr{Cvoid}, stream.handle), C_NULL, out)
    rc == 0 || error("scb_create failed with code $rc (an sm_100 GPU is required; there is no CPU fallback)")
    h = Handle(out[], Base.unsafe_convert(Ptr{Cvoid}, stream.handle))
    finalizer(x -> ccall((:scb_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

const _handle = Ref{Union{Nothing, Handle}}(nothing)
default_handle() = (_handle[] === nothing && (_handle[] = Handle()); _handle[]::Handle)

function check(h::Handle, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:scb_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))
    error(msg)   # ErrorException, like the reference's error("...")
end

# ---- Mesh3D: same fields as src/mesh.jl:19-34 ---------------------------------------------------
mutable struct Mesh3D{T <: AbstractFloat, A <: AbstractArray{T}, B <: AbstractArray{T}}
    grid_size::NTuple{3, Int}
    min_bounds::NTuple{3, T}
    max_bounds::NTuple{3, T}
    delta::NTuple{3, T}
    gamma::T
    total_charge::T
    rho::A
    efield::B
    _workspace::Union{Nothing, Handle}   # the reference keeps FFT plans here; we keep the scb handle
end

_n(m::Mesh3D) = Int64[m.grid_size...]
_f3(t) = Float64[t...]

# Work is enqueued on the handle's stream; CUDA.jl gives every task its own stream, so before each call the handle
# follows the calling task's current stream (scb_set_stream synchronises the old stream once when it changes).
function use_current_stream!(h::Handle)
    s = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
    if s != h.stream
        check(h, ccall((:scb_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, s))
        h.stream = s
    end
    return h
end
function handle(m::Mesh3D)
    m._workspace === nothing && (m._workspace = default_handle())
    return use_current_stream!(m._workspace::Handle)
end

# The reference promotes mixed Float32 / Float64 particle arrays element by element (src/deposition.jl:39-41); the C ABI
# takes ONE particle element type per call, so mixed arrays are refused instead of being read with the wrong size.
function particle_eltype(arrays...)
    P = eltype(arrays[1])
    all(a -> eltype(a) === P, arrays) || error("particle arrays must share one element type (got $(map(eltype, arrays)))")
    (P === Float32 || P === Float64) || error("only Float32 and Float64 particle arrays are supported")
    return P
end

"""
    Mesh3D(grid_size, particles_x, particles_y, particles_z; T=Float64, gamma=1.0, total_charge=0.0)

Bounds from the particle extrema, arithmetic of src/mesh.jl:118-156 (extrema and first delta in the
particles' precision, 1e-6 padding and final delta in Float64, zero delta -> 1e-6, then cast to T).
For `CuArray` particles the extrema come from `scb_bounds` (one fused device reduction).
"""
function Mesh3D(grid_size::NTuple{3, Int}, particles_x, particles_y, particles_z;
                T::Type{<:AbstractFloat} = Float64, backend = nothing, gamma::Real = 1.0, total_charge::Real = 0.0)
    any(grid_size .<= 1) && error("All elements of grid_size must be at least 2.")
    (isempty(particles_x) || isempty(particles_y) || isempty(particles_z)) && error("Particle arrays cannot be empty.")
    (length(particles_x) == length(particles_y) == length(particles_z)) ||
        error("Particle coordinate arrays must have the same length.")
    if particles_x isa CuArray
        P = eltype(particles_x)
        lo = zeros(Float64, 3); hi = zeros(Float64, 3)
        particle_eltype(particles_x, particles_y, particles_z)
        h = use_current_stream!(default_handle())
        check(h, ccall((:scb_bounds, LIB), Cint,
                       (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}),
                       h.ptr, length(particles_x), particles_x, particles_y, particles_z, dtag(P), lo, hi))
        mins = (P(lo[1]), P(lo[2]), P(lo[3])); maxs = (P(hi[1]), P(hi[2]), P(hi[3]))
    else
        ex = (extrema(particles_x), extrema(particles_y), extrema(particles_z))
        mins = (ex[1][1], ex[2][1], ex[3][1]); maxs = (ex[1][2], ex[2][2], ex[3][2])
    end
    delta = ntuple(i -> (maxs[i] - mins[i]) / (grid_size[i] - 1), 3)
    mins = ntuple(i -> mins[i] - 1e-6 * delta[i], 3)
    maxs = ntuple(i -> maxs[i] + 1e-6 * delta[i], 3)
    delta = ntuple(i -> (maxs[i] - mins[i]) / (grid_size[i] - 1), 3)
    delta = ntuple(i -> delta[i] == 0 ? 1e-6 : delta[i], 3)
    rho = CUDA.zeros(T, grid_size...)
    efield = CUDA.zeros(T, grid_size..., 3)
    return Mesh3D{T, typeof(rho), typeof(efield)}(grid_size, T.(mins), T.(maxs), T.(delta), T(gamma), T(total_charge),
                                                  rho, efield, nothing)
end

"""
    Mesh3D(grid_size, min_bounds, max_bounds; T=Float64, gamma=1.0, total_charge=0.0)

Manual bounds (src/mesh.jl:196-238): bounds cast to T first, delta computed in T.
"""
function Mesh3D(grid_size::NTuple{3, Int}, min_bounds::NTuple{3, Real}, max_bounds::NTuple{3, Real};
                T::Type{<:AbstractFloat} = Float64, backend = nothing, gamma::Real = 1.0, total_charge::Real = 0.0)
    any(grid_size .<= 1) && error("All elements of grid_size must be at least 2.")
    any(max_bounds .<= min_bounds) && error("max_bounds must be strictly greater than min_bounds for all dimensions.")
    lo = T.(min_bounds); hi = T.(max_bounds)
    delta = (hi .- lo) ./ (grid_size .- 1)
    rho = CUDA.zeros(T, grid_size...)
    efield = CUDA.zeros(T, grid_size..., 3)
    return Mesh3D{T, typeof(rho), typeof(efield)}(grid_size, lo, hi, delta, T(gamma), T(total_charge), rho, efield, nothing)
end

function Base.show(io::IO, mesh::Mesh3D{T}) where {T}
    nx, ny, nz = mesh.grid_size
    lo, hi = mesh.min_bounds, mesh.max_bounds
    print(io, "Mesh3D{$T, CuArray} ($(nx)x$(ny)x$(nz)) bounds=[($(lo[1]),$(lo[2]),$(lo[3])), ($(hi[1]),$(hi[2]),$(hi[3]))] gamma=$(mesh.gamma)")
end

# ---- clear_mesh!  (src/deposition.jl:10-12) -----------------------------------------------------
function clear_mesh!(mesh::Mesh3D{T}) where {T}
    h = handle(mesh)
    check(h, ccall((:scb_clear, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Ptr{Int64}, Cint),
                   h.ptr, mesh.rho, _n(mesh), dtag(T)))
end

# ---- deposit!  (src/deposition.jl:218-247) -------------------------------------------------------
function deposit!(mesh::Mesh3D{T}, particles_x, particles_y, particles_z, particles_q; clear::Bool = true) where {T}
    (length(particles_x) == length(particles_y) == length(particles_z) == length(particles_q)) ||
        error("Particle coordinate and charge arrays must have the same length.")
    particles_x isa CuArray || error("Unsupported backend: particle arrays must be CuArrays")
    P = particle_eltype(particles_x, particles_y, particles_z, particles_q)
    h = handle(mesh)
    check(h, ccall((:scb_deposit, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid}, Cint,
                    Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cint),
                   h.ptr, length(particles_x), particles_x, particles_y, particles_z, particles_q, dtag(P), mesh.rho,
                   dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.delta), clear ? 1 : 0))
end

# ---- solve! / solve_freespace!  (src/solvers/free_space.jl:14-47, 56-101) -----------------------
function solve!(mesh::Mesh3D{T}; at_cathode::Bool = false) where {T}
    h = handle(mesh)
    check(h, ccall((:scb_solve, LIB), Cint,
                   (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Float64, Cint),
                   h.ptr, mesh.rho, mesh.efield, dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.max_bounds),
                   _f3(mesh.delta), Float64(mesh.gamma), at_cathode ? 1 : 0))
end

function solve_freespace!(mesh::Mesh3D{T}; offset::NTuple{3, T} = (zero(T), zero(T), zero(T))) where {T}
    h = handle(mesh)
    check(h, ccall((:scb_solve_freespace, LIB), Cint,
                   (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Float64, Ptr{Float64}),
                   h.ptr, mesh.rho, mesh.efield, dtag(T), _n(mesh), _f3(mesh.delta), Float64(mesh.gamma), _f3(offset)))
end

# ---- interpolate_field  (src/interpolation.jl:100-128) ------------------------------------------
function interpolate_field(mesh::Mesh3D{T}, particles_x, particles_y, particles_z) where {T}
    P = particle_eltype(particles_x, particles_y, particles_z)
    Ex = similar(particles_x); Ey = similar(particles_x); Ez = similar(particles_x)
    h = handle(mesh)
    check(h, ccall((:scb_interpolate, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid}, Cint, Ptr{Int64},
                    Ptr{Float64}, Ptr{Float64}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}),
                   h.ptr, length(particles_x), particles_x, particles_y, particles_z, dtag(P), mesh.efield, dtag(T),
                   _n(mesh), _f3(mesh.min_bounds), _f3(mesh.delta), Ex, Ey, Ez))
    return Ex, Ey, Ez
end

# ---- get_green_function!  (src/green_functions.jl:41-67), parity hook ---------------------------
"""
Fills `cgrn` (Complex{T}, size (2nx,2ny,2nz)) like the reference: real part = integrated Green
function with raw last planes, imaginary part = 0.
"""
function get_green_function!(cgrn::CuArray{Complex{T}, 3}, delta::NTuple{3, T}, gamma::T, icomp::Int;
                             offset::NTuple{3, T} = (zero(T), zero(T), zero(T)), temp = nothing) where {T}
    re = CUDA.zeros(T, size(cgrn)...)
    h = use_current_stream!(default_handle())
    check(h, ccall((:scb_green, LIB), Cint,
                   (Ptr{Cvoid}, CuPtr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Float64, Cint, Ptr{Float64}, Cint),
                   h.ptr, re, Int64[size(cgrn)...], _f3(delta), Float64(gamma), icomp, _f3(offset), dtag(T)))
    cgrn .= complex.(re)
    return cgrn
end

# src/green_functions.jl:13-22, 35-38 (host-side scalar helpers kept for API completeness)
@inline function field_green_function(x, y, z)
    r = sqrt(x^2 + y^2 + z^2)
    return x * atan((y * z) / (r * x)) - z * log(r + y) + y * log((r - z) / (r + z)) / 2
end

@inline function potential_green_function(x, y, z)
    r = sqrt(x^2 + y^2 + z^2)
    r == zero(x) && return zero(x)
    half = one(x) / 2
    return -half * z^2 * atan(x * y / (z * r)) - half * y^2 * atan(x * z / (y * r)) -
           half * x^2 * atan(y * z / (x * r)) + y * z * log(x + r) + x * z * log(y + r) + x * y * log(z + r)
end

# ---- extensions (not in the reference): potential, magnetic field, fused kick -------------------
"""
    solve_potential!(mesh; at_cathode=false) -> phi::CuArray{T,3}

`solve!` that also returns the scalar potential (rest-frame, `potential_green_function` as a fourth
component of the same fused convolution; scb_solve_potential).  `mesh.efield` is written as by `solve!`.
"""
function solve_potential!(mesh::Mesh3D{T}; at_cathode::Bool = false) where {T}
    phi = similar(mesh.rho)
    h = handle(mesh)
    check(h, ccall((:scb_solve_potential, LIB), Cint,
                   (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Float64, Cint),
                   h.ptr, mesh.rho, mesh.efield, phi, dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.max_bounds),
                   _f3(mesh.delta), Float64(mesh.gamma), at_cathode ? 1 : 0))
    return phi
end

"""
    magnetic_field(mesh) -> B::CuArray{T,4}

B = (beta/c) z_hat x E of a bunch moving along +z with `mesh.gamma` (scb_bfield).
"""
function magnetic_field(mesh::Mesh3D{T}) where {T}
    b = similar(mesh.efield)
    h = handle(mesh)
    check(h, ccall((:scb_bfield, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Int64}, Float64),
                   h.ptr, mesh.efield, b, dtag(T), _n(mesh), Float64(mesh.gamma)))
    return b
end

"""
    interpolate_kick!(mesh, x, y, z, px, py, pz, coef_xy, coef_z)

`interpolate_field` fused with the momentum update `p .+= coef .* E` (scb_interpolate_kick): the
interpolated field never round-trips through memory.
"""
function interpolate_kick!(mesh::Mesh3D{T}, x, y, z, px, py, pz, coef_xy::Real, coef_z::Real) where {T}
    P = particle_eltype(x, y, z, px, py, pz)
    h = handle(mesh)
    check(h, ccall((:scb_interpolate_kick, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid}, Cint, Ptr{Int64},
                    Ptr{Float64}, Ptr{Float64}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Float64),
                   h.ptr, length(x), x, y, z, dtag(P), mesh.efield, dtag(T), _n(mesh), _f3(mesh.min_bounds),
                   _f3(mesh.delta), px, py, pz, Float64(coef_xy), Float64(coef_z)))
end

"""
    step!(mesh, x, y, z, q, Ex, Ey, Ez; at_cathode=false)

deposit! + solve! + interpolate_field in one call with caller-owned outputs (scb_step).
"""
function step!(mesh::Mesh3D{T}, x, y, z, q, Ex, Ey, Ez; at_cathode::Bool = false) where {T}
    P = particle_eltype(x, y, z, q, Ex, Ey, Ez)
    h = handle(mesh)
    check(h, ccall((:scb_step, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid},
                    CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                    CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}),
                   h.ptr, length(x), x, y, z, q, dtag(P), mesh.rho, mesh.efield, dtag(T), _n(mesh),
                   _f3(mesh.min_bounds), _f3(mesh.max_bounds), _f3(mesh.delta), Float64(mesh.gamma), at_cathode ? 1 : 0,
                   Ex, Ey, Ez))
end

"""
    step_host!(mesh, x, y, z, q, Ex, Ey, Ez; at_cathode=false, wait=true)

The same step with HOST arrays (`Array{P}`, ideally pinned with `CUDA.pin`): uploads, deposits, solves, interpolates and
downloads in overlapping chunks (scb_step_host).  With `wait=false` the call returns as soon as everything is queued
(scb_step_host_async): up to two steps may be in flight, the upload of one overlapping the download of the previous one;
consecutive steps need distinct output arrays, and `step_host_wait!(mesh)` must return before any of the arrays is
touched.
"""
function step_host!(mesh::Mesh3D{T}, x::Array{P}, y::Array{P}, z::Array{P}, q::Array{P}, Ex::Array{P}, Ey::Array{P},
                    Ez::Array{P}; at_cathode::Bool = false, wait::Bool = true) where {T,P}
    h = handle(mesh)
    args = (h.ptr, length(x), pointer(x), pointer(y), pointer(z), pointer(q), dtag(P), mesh.rho, mesh.efield, dtag(T),
            _n(mesh), _f3(mesh.min_bounds), _f3(mesh.max_bounds), _f3(mesh.delta), Float64(mesh.gamma),
            at_cathode ? 1 : 0, pointer(Ex), pointer(Ey), pointer(Ez))
    sig = (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint,
           Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid})
    GC.@preserve x y z q Ex Ey Ez begin
        if wait
            check(h, ccall((:scb_step_host, LIB), Cint, sig, args...))
        else
            check(h, ccall((:scb_step_host_async, LIB), Cint, sig, args...))
        end
    end
end

step_host_wait!(mesh::Mesh3D) = (h = handle(mesh); check(h, ccall((:scb_step_host_wait, LIB), Cint, (Ptr{Cvoid},), h.ptr)))

# ---- multi-GPU (extension; the reference is single-device): one Julia process and one handle per GPU ----------
"""
    comm_unique_id() -> Vector{UInt8}          (on one rank; broadcast the 128 bytes, e.g. with MPI.Bcast!)
    comm_init!(mesh, nranks, rank, uid)        (on every rank)

Create the library's NCCL communicator for the mesh's handle (scb_comm_unique_id / scb_comm_init).
"""
function comm_unique_id()
    uid = zeros(UInt8, 128)
    rc = ccall((:scb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), uid)
    rc == 0 || error("scb_comm_unique_id failed with code $rc (libnccl.so.2 not loadable?)")
    return uid
end
comm_init!(mesh::Mesh3D, nranks::Integer, rank::Integer, uid::Vector{UInt8}) =
    (h = handle(mesh); check(h, ccall((:scb_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), h.ptr, nranks, rank, uid)))
comm_destroy!(mesh::Mesh3D) = (h = handle(mesh); check(h, ccall((:scb_comm_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr)))

"""
    solve_sharded!(mesh; at_cathode=false)

`solve!` for a particle-sharded run: `mesh.rho` holds this rank's partial charge grid (every rank deposited its own
shard on the same geometry); slab-decomposed solve over NCCL / NVLink, every rank ends with the full `mesh.efield`
(scb_solve_sharded).
"""
function solve_sharded!(mesh::Mesh3D{T}; at_cathode::Bool = false) where {T}
    h = handle(mesh)
    check(h, ccall((:scb_solve_sharded, LIB), Cint,
                   (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Float64, Cint),
                   h.ptr, mesh.rho, mesh.efield, dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.max_bounds),
                   _f3(mesh.delta), Float64(mesh.gamma), at_cathode ? 1 : 0))
end

"""
    step_sharded!(mesh, x, y, z, q, Ex, Ey, Ez; at_cathode=false)

`step!` on this rank's particle shard (scb_step_sharded).
"""
function step_sharded!(mesh::Mesh3D{T}, x, y, z, q, Ex, Ey, Ez; at_cathode::Bool = false) where {T}
    P = particle_eltype(x, y, z, q, Ex, Ey, Ez)
    h = handle(mesh)
    check(h, ccall((:scb_step_sharded, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid},
                    CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                    CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}),
                   h.ptr, length(x), x, y, z, q, dtag(P), mesh.rho, mesh.efield, dtag(T), _n(mesh),
                   _f3(mesh.min_bounds), _f3(mesh.max_bounds), _f3(mesh.delta), Float64(mesh.gamma), at_cathode ? 1 : 0,
                   Ex, Ey, Ez))
end

"""
    step_host_sharded!(mesh, x, y, z, q, Ex, Ey, Ez; at_cathode=false, slab_solve=true)

`step_host!(...; wait=false)` for one rank of a particle-sharded run (scb_step_host_sharded_async): the host arrays are
this rank's shard, `mesh.rho` receives the rank's partial charge grid and `mesh.efield` the full field.  The handle must
carry a communicator (`comm_init!`); every rank queues the same sequence of steps and finishes with `step_host_wait!`.
"""
function step_host_sharded!(mesh::Mesh3D{T}, x::Array{P}, y::Array{P}, z::Array{P}, q::Array{P}, Ex::Array{P},
                            Ey::Array{P}, Ez::Array{P}; at_cathode::Bool = false, slab_solve::Bool = true) where {T,P}
    h = handle(mesh)
    GC.@preserve x y z q Ex Ey Ez begin
        check(h, ccall((:scb_step_host_sharded_async, LIB), Cint,
                       (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid},
                        Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Cint, Ptr{Cvoid},
                        Ptr{Cvoid}, Ptr{Cvoid}),
                       h.ptr, length(x), pointer(x), pointer(y), pointer(z), pointer(q), dtag(P), mesh.rho, mesh.efield,
                       dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.max_bounds), _f3(mesh.delta), Float64(mesh.gamma),
                       at_cathode ? 1 : 0, slab_solve ? 1 : 0, pointer(Ex), pointer(Ey), pointer(Ez)))
    end
end

# ---- strided / array-of-structures particle records (extension, scb_*_strided) -------------------
# Mirrors `scb_particle_strides` (include/spacecharge_b200.h): ELEMENT strides of the seven particle arrays.
struct ParticleStrides
    x::Int64; y::Int64; z::Int64; q::Int64
    ex::Int64; ey::Int64; ez::Int64
    reserved::Int64
end

"""
    deposit!(mesh, records::CuMatrix, q; rows=(1, 3, 5), clear=true)

`deposit!` on Bmad-style phase-space records: `records` is a `(6, Np)` matrix `(x, px, y, py, z, pz)` per column;
the coordinates are read in place (scb_deposit_strided, element stride `size(records, 1)`).  `q` is a dense charge
vector, or a single number for equal-weight macro-particles (charge stride 0).
"""
function deposit!(mesh::Mesh3D{T}, records::CuMatrix{P}, q; rows::NTuple{3, Int} = (1, 3, 5), clear::Bool = true) where {T, P}
    np = size(records, 2)
    qarr = q isa Number ? CuArray(P[q]) : q
    st = Ref(ParticleStrides(size(records, 1), size(records, 1), size(records, 1), q isa Number ? 0 : 1, 1, 1, 1, 0))
    xs, ys, zs = (pointer(records, r) for r in rows)
    h = handle(mesh)
    GC.@preserve records qarr check(h, ccall((:scb_deposit_strided, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{ParticleStrides}, Cint,
                    CuPtr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cint),
                   h.ptr, np, xs, ys, zs, qarr, st, dtag(P), mesh.rho, dtag(T), _n(mesh), _f3(mesh.min_bounds),
                   _f3(mesh.delta), clear ? 1 : 0))
end

"""
    interpolate_kick!(mesh, records::CuMatrix, coef_xy, coef_z; rows=(1, 3, 5), prows=(2, 4, 6))

Gather fused with the momentum update on the same records (scb_interpolate_kick_strided): rows `rows` are read, rows
`prows` are updated in place, the bunch is never copied or de-interleaved.
"""
function interpolate_kick!(mesh::Mesh3D{T}, records::CuMatrix{P}, coef_xy::Real, coef_z::Real;
                           rows::NTuple{3, Int} = (1, 3, 5), prows::NTuple{3, Int} = (2, 4, 6)) where {T, P}
    s = size(records, 1)
    st = Ref(ParticleStrides(s, s, s, 1, s, s, s, 0))
    xs, ys, zs = (pointer(records, r) for r in rows)
    pxs, pys, pzs = (pointer(records, r) for r in prows)
    h = handle(mesh)
    GC.@preserve records check(h, ccall((:scb_interpolate_kick_strided, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{ParticleStrides}, Cint, CuPtr{Cvoid}, Cint,
                    Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Float64),
                   h.ptr, size(records, 2), xs, ys, zs, st, dtag(P), mesh.efield, dtag(T), _n(mesh), _f3(mesh.min_bounds),
                   _f3(mesh.delta), pxs, pys, pzs, Float64(coef_xy), Float64(coef_z)))
end

"""
    interpolate_field(mesh, records::CuMatrix; rows=(1, 3, 5)) -> (Ex, Ey, Ez)

`interpolate_field` with the coordinates read from the records in place (scb_interpolate_strided); the outputs are
fresh dense vectors like the reference's.
"""
function interpolate_field(mesh::Mesh3D{T}, records::CuMatrix{P}; rows::NTuple{3, Int} = (1, 3, 5)) where {T, P}
    np = size(records, 2)
    s = size(records, 1)
    Ex = CuArray{P}(undef, np); Ey = similar(Ex); Ez = similar(Ex)
    st = Ref(ParticleStrides(s, s, s, 1, 1, 1, 1, 0))
    xs, ys, zs = (pointer(records, r) for r in rows)
    h = handle(mesh)
    GC.@preserve records check(h, ccall((:scb_interpolate_strided, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{ParticleStrides}, Cint, CuPtr{Cvoid}, Cint,
                    Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}),
                   h.ptr, np, xs, ys, zs, st, dtag(P), mesh.efield, dtag(T), _n(mesh), _f3(mesh.min_bounds),
                   _f3(mesh.delta), Ex, Ey, Ez))
    return Ex, Ey, Ez
end

# ---- bunches kept ordered by cell (extension; scb_sort_particles, scb_permute, scb_set_particle_order) -------------
const SCB_ORDER_RANDOM = Cint(0)
const SCB_ORDER_CELL = Cint(1)
const SCB_ORDER_AUTO = Cint(3)

"""
    sort_particles(mesh, x, y, z) -> perm::CuVector{UInt32}

Permutation that orders the bunch by linear cell index of `mesh` (0-based particle indices, stable).
"""
function sort_particles(mesh::Mesh3D{T}, x::CuVector, y::CuVector, z::CuVector) where {T}
    P = particle_eltype(x, y, z)
    (length(x) == length(y) == length(z)) || error("Particle coordinate arrays must have the same length.")
    perm = CuVector{UInt32}(undef, length(x))
    h = handle(mesh)
    check(h, ccall((:scb_sort_particles, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Float64},
                    Ptr{Float64}, CuPtr{Cvoid}),
                   h.ptr, length(x), x, y, z, dtag(P), dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.delta), perm))
    return perm
end

"""
    permute(mesh, perm, arrays...) -> Tuple of new CuVectors, `a[perm .+ 1]` for every array (up to 8 per pass)
"""
function permute(mesh::Mesh3D, perm::CuVector{UInt32}, arrays::CuVector...)
    P = particle_eltype(arrays...)
    all(a -> length(a) == length(perm), arrays) || error("permute: arrays must have the length of perm")
    outs = map(similar, arrays)
    h = handle(mesh)
    for first in 1:8:length(arrays)
        last = min(first + 7, length(arrays))
        src = [reinterpret(Ptr{Cvoid}, pointer(a)) for a in arrays[first:last]]
        dst = [reinterpret(Ptr{Cvoid}, pointer(a)) for a in outs[first:last]]
        GC.@preserve arrays outs check(h, ccall((:scb_permute, LIB), Cint,
                       (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint),
                       h.ptr, length(perm), perm, last - first + 1, src, dst, dtag(P)))
    end
    return outs
end

"""
    sort_particles!(mesh, x, y, z, others...) -> (perm, x_sorted, y_sorted, z_sorted, others_sorted...)
"""
function sort_particles!(mesh::Mesh3D, x::CuVector, y::CuVector, z::CuVector, others::CuVector...)
    perm = sort_particles(mesh, x, y, z)
    return (perm, permute(mesh, perm, x, y, z, others...)...)
end

"""
    set_particle_order!(mesh, order)    order = :random (default kernels), :cell (run-accumulating kernels) or :auto
                                        (the handle samples the order every eighth deposit and picks)

Results do not depend on the setting; only the speed of `deposit!`, `interpolate_field` and `step!` does.
"""
function set_particle_order!(mesh::Mesh3D, order::Symbol)
    code = order === :cell ? SCB_ORDER_CELL : order === :random ? SCB_ORDER_RANDOM : order === :auto ? SCB_ORDER_AUTO :
           error("order must be :random, :cell or :auto")
    h = handle(mesh)
    check(h, ccall((:scb_set_particle_order, LIB), Cint, (Ptr{Cvoid}, Cint), h.ptr, code))
end

"""
    particle_order_fraction(mesh, x, y, z) -> Float64

Fraction of sampled neighbouring particle pairs in the same or the x-adjacent cell (1 = ordered, 0 = random).
"""
function particle_order_fraction(mesh::Mesh3D{T}, x::CuVector, y::CuVector, z::CuVector) where {T}
    P = particle_eltype(x, y, z)
    out = Ref{Float64}(0.0)
    h = handle(mesh)
    check(h, ccall((:scb_particle_order_fraction, LIB), Cint,
                   (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Float64},
                    Ptr{Float64}, Ref{Float64}),
                   h.ptr, length(x), x, y, z, dtag(P), dtag(T), _n(mesh), _f3(mesh.min_bounds), _f3(mesh.delta), out))
    return out[]
end

"""
    allreduce_rho!(mesh)

Sum of the ranks' charge grids in place on the library's communicator (scb_allreduce_rho): the "solve replicated" mode
of a particle-sharded run.
"""
allreduce_rho!(mesh::Mesh3D{T}) where {T} =
    (h = handle(mesh); check(h, ccall((:scb_allreduce_rho, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Ptr{Int64}, Cint),
                                      h.ptr, mesh.rho, _n(mesh), dtag(T))))

end # module SpaceChargeB200
