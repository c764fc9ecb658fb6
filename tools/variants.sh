#!/bin/bash
# usage: tools/variants.sh  -- solver A/B (config 2, both state dtypes) for every tuning build of the CUDA library under
# build/variants.  The product loader takes no environment override: tools/sor_ab.py --library PATH installs the
# tuning build for that one process (still CUDA: _lib._select_for_tests(path, emulator=False)).
for so in base build/variants/*.so; do
  if [ "$so" = base ]; then LIBARG=""; else LIBARG="--library $PWD/$so"; fi
  echo "== $so"; python tools/sor_ab.py $LIBARG "$@" 2>&1 | cut -c1-260
done
