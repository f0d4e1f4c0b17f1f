# record_replay.jl -- for users WITH Julia: run the REAL Arianna reference on `particle_1d` while recording every
# random draw, and dump (u_cat, z, u_acc), the accept/reject decisions and the trajectory so that
# `arianna_sweep_replay` (replay mode) can be compared bit for bit against the reference itself.
# NOT executed in the build environment (no Julia toolchain).
#
#   julia --project=. julia/tools/record_replay.jl <M> <steps> <out.bin>
using Arianna, Random, Distributions, ComponentArrays
include(joinpath(dirname(pathof(Arianna)), "..", "example", "particle_1d", "particle_1d.jl"))
potential(x) = x^2

"An AbstractRNG that forwards to a real Xoshiro and logs what the hot path draws."
mutable struct RecordingRNG <: AbstractRNG
    inner::Xoshiro
    uniforms::Vector{Float64}
    normals::Vector{Float64}
end
RecordingRNG(seed) = RecordingRNG(Xoshiro(seed), Float64[], Float64[])
Random.rand(r::RecordingRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = (u = rand(r.inner); push!(r.uniforms, u); u)
Random.randn(r::RecordingRNG, ::Type{Float64}) = (z = randn(r.inner); push!(r.normals, z); z)

function main(M, steps, out)
    seed, β, σ = 42, 2.0, 0.1
    rng0 = Xoshiro(seed)
    x0 = [4rand(rng0) - 2 for _ in 1:M]
    open(out, "w") do io
        write(io, Int64(M), Int64(steps), β, σ)
        write(io, x0)
        for c in 1:M
            sys = System(x0[c], β)
            pool = (Move(Displacement(0.0), StandardGaussian(), ComponentArray(σ=σ), 1.0),)
            rng = RecordingRNG(seed + c - 1)                      # metropolis.jl:262
            dec = UInt8[]
            for _ in 1:steps
                before = pool[1].accepted_calls
                mc_sweep!(sys, pool, rng; mc_steps=1)             # u_cat, z, u_acc in this order
                push!(dec, UInt8(pool[1].accepted_calls - before))
            end
            # per chain: u_cat[steps], z[steps], u_acc[steps], decisions[steps], final x
            write(io, rng.uniforms[1:2:end], rng.normals, rng.uniforms[2:2:end], dec, sys.x)
        end
    end
end
# Seeding check for montecarlo_b200/julia_rng.py and oracle/arianna_oracle.c:ao_xoshiro_seed_julia (SHA-256 of the seed's
# 32-bit limbs, Julia 1.7 - 1.10): they compute Xoshiro(42) = (a379de7eeeb2a4e8, 953dccb6b532b3af, f597b8ff8cfd652a,
# ccd7337c571680d1); a different line here means this Julia version seeds differently (1.11+) and the device xoshiro
# mode must be fed uploaded states (CudaEnsemble(...; rng=:xoshiro) of the shim does exactly that).
let r = Xoshiro(42)
    println("Xoshiro(42) state: ", join(string.((r.s0, r.s1, r.s2, r.s3), base=16), " "))
end
main(parse(Int, ARGS[1]), parse(Int, ARGS[2]), ARGS[3])
