#!/bin/bash
# copy the working copy under .wip/ (where development happens while a gpurun call may be snapshotting the tree)
# over the repo root, including the built objects so `make` at the root is incremental
cd /tmp/wip || exit 1
find . -type d \( -name __pycache__ -o -name .pytest_cache \) -prune -o -type f -print | while read -r f; do
  mkdir -p "/root/repo/$(dirname "$f")"
  cp -p "$f" "/root/repo/$f"
done
