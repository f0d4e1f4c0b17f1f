# check_prediction.jl -- for users WITH Julia (1.7 - 1.10): run the reference's own example script
# (example/particle_1d/harmonic_oscillator/MC_harmonic_oscillator.jl: seed 42, β = 2, M = 10, 10^5 steps) and compare
# what it writes with the prediction committed in tests/golden/julia_prediction_config1.json (computed by the CPU
# oracle in "Julia mode": Julia's Xoshiro seeding, rand, randn and Arianna's mc_step! restated in C).
# Equality closes the "parity unpinned" gap for config 1; the first differing line says where the restatement is off.
# NOT executed in the build environment (no Julia toolchain).
#
#   julia --project=. julia/tools/check_prediction.jl <path/to/Arianna.jl> <path/to/julia_prediction_config1.json>
using SHA

arianna, prediction = ARGS[1], read(ARGS[2], String)
example = joinpath(arianna, "example", "particle_1d", "harmonic_oscillator")
cd(example) do
    # the script ends with a plotting section: run only the part up to run!(simulation)
    src = read("MC_harmonic_oscillator.jl", String)
    cut = findfirst("## PLOT RESULTS", src)
    include_string(Main, "using Arianna, Random, ComponentArrays\n" * src[1:first(cut)-1])
end
out = joinpath(example, "data", "MC", "particle_1d", "Harmonic", "beta2.0", "M10", "seed42")

# a JSON reader for the three kinds of value the prediction holds would be a dependency; the file is one `"key": value`
# per line, so plain string search is enough
field(key) = match(Regex("\"$key\": \"([0-9a-f]+)\""), prediction).captures[1]
function head(key)
    m = match(Regex("\"$key\": \\[(.*?)\\]\\s*,\\s*\"", "s"), prediction)
    [String(x.captures[1]) for x in eachmatch(r"\"([^\"]*)\"", m.captures[1])]
end
ok = true
for name in ("energy", "acceptance")
    text = read(joinpath(out, "$name.dat"), String)
    same = bytes2hex(sha256(text)) == field("$(name)_sha256")
    println("$name.dat: ", same ? "IDENTICAL to the prediction (sha256)" : "differs")
    if !same
        global ok = false
        for (i, (a, b)) in enumerate(zip(split(text, "\n"), head("$(name)_head")))
            a == b || (println("  first difference within the committed head at line $i:\n    reference : $a\n    prediction: $b"); break)
        end
    end
end
exit(ok ? 0 : 1)
