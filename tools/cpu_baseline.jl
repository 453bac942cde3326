# cpu_baseline.jl -- times the UNMODIFIED reference (BioDemuX.jl) on config 2's workload, for anyone with Julia.
# NOT RUN in this repository (no Julia in the build image): bench.py's cpu_baseline / --impl reference legs time the
# C restatement of the same algorithm instead and say so.
#
#   python tools/make_config2_fastq.py /tmp/cfg2 1000000
#   JULIA_NUM_THREADS=$(nproc) julia --project=/path/to/BioDemuX.jl tools/cpu_baseline.jl /tmp/cfg2
#
# Prints reads/s and full-matrix GCUPS (96 barcodes x 24 nt x 150 columns per read, SURVEY.md section 8d).
using BioDemuX

dir = ARGS[1]
fastq, bcs, out = joinpath(dir, "reads.fastq"), joinpath(dir, "barcodes.csv"), joinpath(dir, "out_julia")
n_reads = countlines(fastq) ÷ 4
rm(out; force=true, recursive=true)
execute_demultiplexing(fastq, bcs, out)                      # warm-up (compilation)
rm(out; force=true, recursive=true)
t = @elapsed execute_demultiplexing(fastq, bcs, out)
println("threads = ", Threads.nthreads(), "  reads = ", n_reads, "  seconds = ", round(t, digits=2))
println("reads/s = ", round(n_reads / t), "   GCUPS (full matrix) = ", round(n_reads * 96 * 24 * 150 / t / 1e9, digits=2))
