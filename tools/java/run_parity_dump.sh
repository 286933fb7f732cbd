#!/bin/bash
# Pins the oracle against the real reference: compiles the headless Starfish sources + ParityDump with javac and runs
# every exported golden case.  Needs a JDK >= 11 and a checkout of particleincell/Starfish.
#   tools/java/run_parity_dump.sh /path/to/Starfish
set -euo pipefail
REF=${1:?path to the Starfish checkout}
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
python "$HERE/export_cases.py"
BUILD=$(mktemp -d)
# headless build: everything but the GUI (buildHeadless.sh of the reference does the same)
find "$REF/src" -name '*.java' ! -path '*/gui/*' ! -name 'Main.java' > "$BUILD/sources.txt"
echo "$HERE/ParityDump.java" >> "$BUILD/sources.txt"
javac -nowarn -d "$BUILD/classes" @"$BUILD/sources.txt"
for d in "$ROOT"/tests/golden/java/*/; do
  name=$(basename "$d")
  java -cp "$BUILD/classes" starfish.core.materials.ParityDump "$d"
  mv "$d/out.txt" "$ROOT/tests/golden/java/$name.txt"
  echo "pinned $name"
done
echo "now: python -m pytest tests/test_golden.py -k java_reference   and commit tests/golden/java/*.txt"
