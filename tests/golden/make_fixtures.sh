#!/bin/sh
# Regenerates tests/golden/reference_test_data.tar.xz from the reference's own
# test DATA (inputs, barcode tables and golden outputs -- no source code):
#   test/FASTQ_files   96 FASTQ inputs (demo1 plain, demo2 gzip)
#   test/reference_files   barcode tables demo1.tsv / demo2.csv (+ unused twins)
#   test/results       124 golden output files of test/integration/single_barcode.jl:2-45
# Run in the build container, where /root/reference is mounted.
set -e
cd "$(dirname "$0")"
tar -C /root/reference/test -cf - FASTQ_files reference_files results | xz -9e > reference_test_data.tar.xz
