# make_golden.jl -- golden vectors FROM THE UNMODIFIED REFERENCE for the regimes its own tests leave unpinned
# (SURVEY.md section 8c "gaps"): traceback through indels, barcode_start_range / barcode_end_range with errors
# allowed, sub-ranges, variable-length sets with a running threshold, nindel != indel, match != 0, m > n.
# NOT RUN in this repository (no Julia in the build image).  Anyone with Julia >= 1.10 and BioDemuX.jl:
#
#   julia --project=/path/to/BioDemuX.jl tools/make_golden.jl tests/golden/julia_vectors.jsonl [n_cases]
#
# writes one JSON object per line; tests/test_julia_vectors.py replays the file through the C oracle and the
# Python transcription when it is present (and through the CUDA path in the -m gpu run).
# Inputs come from a small LCG (no dependence on Random's stream across Julia versions).
using BioDemuX
const B = BioDemuX

mutable struct LCG
    s::UInt64
end
next!(g::LCG) = (g.s = g.s * 6364136223846793005 + 1442695040888963407; (g.s >> 33) % Int)
below!(g::LCG, n::Int) = next!(g) % n
pick!(g::LCG, v) = v[below!(g, length(v)) + 1]
randseq!(g::LCG, lo::Int, hi::Int, alpha::String="ACGT") = String([alpha[below!(g, length(alpha)) + 1] for _ in 1:(lo + below!(g, hi - lo + 1))])

function planted!(g::LCG, q::String, nlo::Int, nhi::Int)
    r = collect(randseq!(g, nlo, nhi))
    if below!(g, 10) < 8 && !isempty(r)
        mq = collect(replace(q, "N" => "A"))
        for _ in 1:pick!(g, [0, 0, 1, 1, 2, 3])
            k = below!(g, 3)
            if k == 0 && !isempty(mq)
                mq[below!(g, length(mq)) + 1] = "ACGT"[below!(g, 4) + 1]
            elseif k == 1
                insert!(mq, below!(g, length(mq) + 1) + 1, "ACGT"[below!(g, 4) + 1])
            elseif length(mq) > 1
                deleteat!(mq, below!(g, length(mq)) + 1)
            end
        end
        st = below!(g, length(r)) + 1
        for (k, c) in enumerate(mq)
            st + k - 1 <= length(r) && (r[st + k - 1] = c)
        end
    end
    return String(r)
end

jnum(x::Float64) = isinf(x) ? "\"Inf\"" : (isnan(x) ? "\"NaN\"" : repr(x))
jnum(x::Int) = string(x)
jopt(x) = isnothing(x) ? "null" : string(x)
jstr(s::String) = "\"" * s * "\""
jlist(v::Vector{String}) = "[" * join(jstr.(v), ",") * "]"
jlist(v::Vector{Int}) = "[" * join(string.(v), ",") * "]"

const THR = [0.0, 0.1, 0.2, 0.25, 0.29, 3 / 11, 0.34, 0.4, 0.5, 0.6, 15 / 22, 1.0]

function main()
    path = ARGS[1]
    n_cases = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 20000
    g = LCG(0x42444d58)
    open(path, "w") do io
        for case in 1:n_cases
            if case % 4 != 0
                # ---- one alignment: semiglobal_alignment / semiglobal_alignment_N (classification.jl:447-477)
                nscoring = below!(g, 5) == 0
                q = randseq!(g, 2, 14, nscoring ? "ACGTN" : "ACGT")
                long_q = below!(g, 8) == 0
                r = long_q ? planted!(g, q[1:min(length(q), 4)], 1, 10) : planted!(g, q, 4, 28)
                n = length(r)
                mt = pick!(g, [0, 0, 0, 1, -1])
                mm = pick!(g, [1, 1, 2, 3])
                ind = pick!(g, [1, 1, 2, 3])
                nind = nscoring ? pick!(g, [1, 2, 3]) : nothing
                norm = nscoring ? max(count(!=('N'), q), 1) : length(q)
                lo = 1 + below!(g, max(n ÷ 2, 1))
                hi = lo + below!(g, n - lo + 1)
                max_start = below!(g, 2) == 0 ? n : lo + below!(g, min(8, n - lo) + 1)
                min_end = below!(g, 2) == 0 ? 1 : max(hi - below!(g, 9), 1)
                tb = below!(g, 2) == 0
                trim = tb ? pick!(g, [nothing, 3, 5]) : nothing
                thr = pick!(g, THR)
                ws = B.SemiGlobalWorkspace(length(q), true)
                res = nscoring ?
                      B.semiglobal_alignment_N(ws, q, r, thr, mt, mm, ind, nind, lo:hi, max_start, min_end, norm, trim, tb) :
                      B.semiglobal_alignment(ws, q, r, thr, mt, mm, ind, lo:hi, max_start, min_end, trim, tb)
                score, s, e = res isa Tuple ? res : (res, -1, -1)
                println(io, "{\"kind\":\"align\",\"q\":", jstr(q), ",\"r\":", jstr(r), ",\"max_error\":", jnum(thr),
                        ",\"match\":", mt, ",\"mismatch\":", mm, ",\"indel\":", ind, ",\"nindel\":", jopt(nind),
                        ",\"lo\":", lo, ",\"hi\":", hi, ",\"max_start\":", max_start, ",\"min_end\":", min_end,
                        ",\"norm\":", norm, ",\"traceback\":", tb, ",\"trim\":", jopt(trim),
                        ",\"score\":", jnum(Float64(score)), ",\"start\":", s, ",\"end\":", e, "}")
            else
                # ---- find_best_matching_bc over a variable-length set (classification.jl:722-728)
                nb = 2 + below!(g, 6)
                bcs = [randseq!(g, 3, 12) for _ in 1:nb]
                below!(g, 3) == 0 && (bcs[below!(g, nb) + 1] = bcs[below!(g, nb) + 1])
                norms = length.(bcs)
                read = planted!(g, pick!(g, bcs), 8, 30)
                n = length(read)
                lo = 1 + below!(g, 4)
                hi = max(lo, n - below!(g, 7))
                max_start = lo + below!(g, min(9, n - lo) + 1)
                min_end = 1 + below!(g, hi)
                md = pick!(g, [0.0, 0.0, 0.05, 0.1, 0.2, 0.34])
                trim = pick!(g, [nothing, nothing, 3, 5])
                need_tb = below!(g, 3) == 0
                cfg = DemuxConfig(max_error_rate=pick!(g, THR), min_delta=md, mismatch=pick!(g, [1, 1, 2]),
                                  indel=pick!(g, [1, 1, 2]), bc_seqs=bcs, bc_lengths_no_N=norms, ids=string.(1:nb),
                                  trim_side=trim)
                ws = B.SemiGlobalWorkspace(maximum(norms), true)
                bc, score, delta, s, e = find_best_matching_bc(read, bcs, norms, cfg, ws, lo:hi, max_start, min_end, trim, need_tb)
                println(io, "{\"kind\":\"find_best\",\"read\":", jstr(read), ",\"barcodes\":", jlist(bcs), ",\"norms\":", jlist(norms),
                        ",\"max_error_rate\":", jnum(cfg.max_error_rate), ",\"min_delta\":", jnum(md),
                        ",\"mismatch\":", cfg.mismatch, ",\"indel\":", cfg.indel, ",\"lo\":", lo, ",\"hi\":", hi,
                        ",\"max_start\":", max_start, ",\"min_end\":", min_end, ",\"trim\":", jopt(trim),
                        ",\"need_tb\":", need_tb, ",\"bc\":", bc, ",\"score\":", jnum(Float64(score)),
                        ",\"delta\":", jnum(Float64(delta)), ",\"start\":", s, ",\"end\":", e, "}")
            end
        end
    end
    println("wrote ", n_cases, " vectors to ", path)
end

main()
