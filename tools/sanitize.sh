#!/bin/bash
# compute-sanitizer evidence (run under gpurun on one B200): memcheck + racecheck on the PRODUCTION build, then all three
# tools on the --synccheck build (the compositor's barrier helpers out of line, see csrc/composite.cu), then the production
# build is restored.  Output: gpurun_out/<tag>_sanitizer.txt
TAG=${1:-r2}
mkdir -p gpurun_out
run() { echo "\$ compute-sanitizer --tool $1 --error-exitcode 9 python tools/sanitize_frame.py   [$2 build]"
        if [ "$1" = racecheck ]; then export B200GS_SANITIZE_WIDE=0; else export B200GS_SANITIZE_WIDE=1; fi   # (see sanitize_frame.py)
        timeout 600 compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 3 python tools/sanitize_frame.py 2>&1 | grep -v "Host Frame\|Saved host" | tail -5
        echo "exit code ${PIPESTATUS[0]}"; echo; }
{
python wgpu-3dgs-viewer-app_b200/build.py --force > /dev/null
run memcheck production; run racecheck production
python wgpu-3dgs-viewer-app_b200/build.py --synccheck > /dev/null
run memcheck synccheck; run racecheck synccheck; run synccheck synccheck
python wgpu-3dgs-viewer-app_b200/build.py --force > /dev/null
} > gpurun_out/${TAG}_sanitizer.txt 2>&1
cat gpurun_out/${TAG}_sanitizer.txt
