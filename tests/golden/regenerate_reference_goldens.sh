#!/bin/bash
# Regenerates every fixture that comes from the reference's own Python code (needs the reference tree: KLAMPT_REFERENCE, default
# /root/reference).  The outputs are deterministic: a clean `git status` afterwards means the committed fixtures are current.
set -e
cd "$(dirname "$0")/../.."
for s in so3 fk mask cspace groupiter iterators rob loader; do
  python tests/golden/make_reference_$s.py > /dev/null 2>&1 || { echo "make_reference_$s.py failed"; exit 1; }
done
git status --short tests/golden
