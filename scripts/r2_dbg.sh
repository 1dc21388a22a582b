TMK_DEBUG_IMAGE=1 python scripts/prof_step.py configs 1 16 2>&1 | grep -i "tmkcl" | head -8
TMK_DEBUG_IMAGE=1 python scripts/prof_step.py blocks 1 16 2>&1 | grep -i "tmkcl" | head -8
TMK_DEBUG_IMAGE=1 python scripts/prof_step.py clutter 1 16 2>&1 | grep -i "tmkcl" | head -8
bash scripts/r2_probe.sh r2f "" "configs clutter" "TMK_WAVE_CHUNK=1048576|TMK_WAVE_CHUNK=524288|TMK_WAVE_CHUNK=262144|TMK_WAVE_CHUNK=2097152"
