#!/bin/bash
# Type-check the OpenRAVE plugin (plugins/cudacollision/*.cpp) against the REAL reference headers in an image that has neither
# Boost nor an OpenRAVE build:  g++ -fsyntax-only with
#   - /root/reference/include            the unmodified public headers (read where they lie)
#   - tools/shim_stubs/                  minimal Boost stand-ins mapped onto the C++17 standard library
#   - a scratch directory                what cmake would generate: openrave/config.h from config.h.in (configure_file) and
#                                        openrave/interfacehashes.h (cpp-gen-md5; any string does for a syntax check)
#   - rapidjson                          from wherever the image has a copy (it ships inside a site-packages tree here)
# A real build links the plugin against the OpenRAVE install (plugins/cudacollision/CMakeLists.txt).
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
[ -d "$REF/include/openrave" ] || { echo "shim_syntax_check: $REF/include/openrave not found (run in the build container)"; exit 2; }
RJ=${RAPIDJSON_INCLUDE:-$(dirname "$(find / -type f -path '*rapidjson/document.h' -not -path '/proc/*' 2>/dev/null | head -1)" 2>/dev/null | xargs -r dirname)}
[ -n "$RJ" ] && [ -f "$RJ/rapidjson/document.h" ] || { echo "shim_syntax_check: no rapidjson headers in this image"; exit 2; }
GEN=$(mktemp -d)
trap 'rm -rf "$GEN"' EXIT
mkdir -p "$GEN/openrave"
python3 - "$REF/config.h.in" "$GEN/openrave/config.h" <<'PY'
import re, sys
s = open(sys.argv[1]).read()
vals = {"OPENRAVE_VERSION_MAJOR": "0", "OPENRAVE_VERSION_MINOR": "149", "OPENRAVE_VERSION_PATCH": "1", "OPENRAVE_PRECISION": "1",
        "OPENRAVE_LOG4CXX": "0", "OPENRAVE_MSGPACK": "0", "OPENRAVE_CURL": "0", "OPENRAVE_ENVIRONMENT_RECURSIVE_LOCK": "1"}
s = re.sub(r"@([A-Za-z0-9_]+)@", lambda m: vals.get(m.group(1), ""), s)
open(sys.argv[2], "w").write(s)
PY
{
  echo "#pragma once"
  for h in COLLISIONCHECKER ROBOT PLANNER KINBODY SENSORSYSTEM CONTROLLER MODULE IKSOLVER PHYSICSENGINE SENSOR TRAJECTORY VIEWER SPACESAMPLER ENVIRONMENT PLUGININFO; do
    echo "#define OPENRAVE_${h}_HASH \"syntax-check\""
  done
} > "$GEN/openrave/interfacehashes.h"
rc=0
for src in "$ROOT"/plugins/cudacollision/*.cpp; do
  echo "== g++ -fsyntax-only $(basename "$src")"
  g++ -std=c++17 -fsyntax-only -Wall -Wno-deprecated-declarations -Wno-unused-local-typedefs \
      -I"$ROOT/tools/shim_stubs" -I"$GEN" -I"$REF/include" -I"$RJ" -I"$ROOT/include" "$src" || rc=1
done
[ $rc -eq 0 ] && echo "shim syntax check: OK"
exit $rc
