#!/bin/sh
# Completes the package skeleton with the reference's own R sources, documentation, data and tests — referenced, not
# copied into this repository:   sh tools/fetch_reference_R.sh /path/to/checkout/of/duckmayr/gpirt
# (everything under R/ except RcppExports.R, whose three lines this skeleton ships itself; man/, data/, tests/ as they are)
set -e
REF="${1:?usage: fetch_reference_R.sh /path/to/duckmayr-gpirt}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
for f in "$REF"/R/*.R; do
    case "$(basename "$f")" in RcppExports.R) ;; *) cp "$f" "$HERE/R/";; esac
done
mkdir -p "$HERE/man" "$HERE/data" "$HERE/tests"
cp -r "$REF"/man/. "$HERE/man/"
cp -r "$REF"/data/. "$HERE/data/"
cp -r "$REF"/tests/. "$HERE/tests/"
# the reference's package documentation file declares useDynLib / importFrom(Rcpp, ...) for roxygen; the NAMESPACE of this
# skeleton is authoritative (no Rcpp), so roxygen is not re-run.
echo "fetched R/, man/, data/, tests/ from $REF; now: R CMD INSTALL $HERE"
