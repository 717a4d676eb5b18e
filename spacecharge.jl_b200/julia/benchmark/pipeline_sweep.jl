# Sweep of the reference's benchmark configurations through the B200 library.
#
#   julia --project pipeline_sweep.jl [deposit|solve|pipeline|all] [--f32] [--compare]
#
# Same workloads as the reference's three benchmark scripts (benchmark/deposit_benchmark.jl,
# benchmark/solve_benchmark.jl, benchmark/full_pipeline_benchmark.jl: six (grid, particles) pairs, Gaussian bunch
# sigma = 1e-3 m, Q = 1e-9 C, Random.seed!(42), minimum over repetitions like @belapsed, mesh construction and
# host->device copies outside the timed region).  `--compare` also times SpaceCharge.jl's own CPU path when that
# package is installed next to this one, so that a maintainer gets the reference's table with one more column.
# Julia is not part of the build image of this repository: the Python twin tools/benchmark_sweep.py is what the
# committed numbers come from; this file is the same driver for users of the Julia shim.
include(joinpath(@__DIR__, "..", "SpaceChargeB200.jl"))
using .SpaceChargeB200
using CUDA
using Random
using Printf

const CONFIGS = [((32, 32, 32), 10_000), ((32, 32, 32), 100_000), ((64, 64, 64), 100_000),
                 ((64, 64, 64), 1_000_000), ((128, 128, 128), 100_000), ((128, 128, 128), 1_000_000)]

function bunch(n, ::Type{P}; sigma = 1.0e-3, total_charge = 1.0e-9) where {P}
    Random.seed!(42)
    x, y, z = (P.(randn(n) .* sigma) for _ in 1:3)
    return x, y, z, fill(P(total_charge / n), n)
end

# minimum wall time of `f` over `reps` runs, each closed by a device synchronisation
function best_of(f, reps)
    f(); CUDA.synchronize()
    best = Inf
    for _ in 1:reps
        t = @elapsed begin
            f(); CUDA.synchronize()
        end
        best = min(best, t)
    end
    return best
end

function time_b200(kind, grid, n, ::Type{T}; reps = 10) where {T}
    x, y, z, q = bunch(n, T)
    dx, dy, dz, dq = CuArray(x), CuArray(y), CuArray(z), CuArray(q)
    mesh = Mesh3D(grid, dx, dy, dz; T = T, total_charge = 1.0e-9)
    deposit!(mesh, dx, dy, dz, dq)
    if kind == :deposit
        return best_of(() -> deposit!(mesh, dx, dy, dz, dq), reps)
    elseif kind == :solve
        return best_of(() -> solve!(mesh), reps)
    end
    return best_of(reps) do
        deposit!(mesh, dx, dy, dz, dq)
        solve!(mesh)
        interpolate_field(mesh, dx, dy, dz)
    end
end

# the reference's CPU path, if SpaceCharge.jl can be loaded (optional column).  The package is loaded at run time, so
# its functions are called through invokelatest (world age).
function time_reference_cpu(kind, grid, n; reps = 3)
    ref = try
        Base.require(Main, :SpaceCharge)
    catch
        return NaN
    end
    il(f, a...; k...) = Base.invokelatest(f, a...; k...)
    x, y, z, q = bunch(n, Float64)
    mesh = il(ref.Mesh3D, grid, x, y, z; total_charge = 1.0e-9)
    il(ref.deposit!, mesh, x, y, z, q)
    run = kind == :deposit ? () -> il(ref.deposit!, mesh, x, y, z, q) :
          kind == :solve ? () -> il(ref.solve!, mesh) :
          () -> (il(ref.deposit!, mesh, x, y, z, q); il(ref.solve!, mesh); il(ref.interpolate_field, mesh, x, y, z))
    run()
    return minimum(@elapsed(run()) for _ in 1:reps)
end

function sweep(kind; T = Float64, compare = false)
    println("\n", kind, "  (", T, ")")
    println("grid    particles    B200 ms", compare ? "    reference CPU ms    ratio" : "")
    for (grid, n) in CONFIGS
        t = time_b200(kind, grid, n, T)
        if compare
            c = time_reference_cpu(kind, grid, n)
            @printf("%-7s %-12d %-10.3f %-19.2f %.1fx\n", "$(grid[1])^3", n, 1e3 * t, 1e3 * c, c / t)
        else
            @printf("%-7s %-12d %-10.3f\n", "$(grid[1])^3", n, 1e3 * t)
        end
    end
end

function main(args)
    T = "--f32" in args ? Float32 : Float64
    compare = "--compare" in args
    kinds = filter(a -> !startswith(a, "--"), args)
    which = isempty(kinds) || kinds[1] == "all" ? [:deposit, :solve, :pipeline] : [Symbol(kinds[1])]
    all(k -> k in (:deposit, :solve, :pipeline), which) ||
        (println("usage: pipeline_sweep.jl [deposit|solve|pipeline|all] [--f32] [--compare]"); exit(1))
    for k in which
        sweep(k; T = T, compare = compare)
    end
end

main(ARGS)
