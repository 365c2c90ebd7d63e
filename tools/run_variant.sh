#!/bin/sh
# usage: tools/run_variant.sh <variant name> <script> [args...] — runs a tool against variants/lib_<name>.so
v=$1; shift; script=$1; shift
python - "$@" <<PY
import os, sys
sys.path.insert(0, ".")
import selfplay_b200.engine as E
E._LIB = os.path.abspath("variants/lib_$v.so")
sys.argv = ["$script"] + sys.argv[1:]
exec(open("$script").read())
PY
