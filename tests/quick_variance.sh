# run-to-run scatter of the house-scene frame time through the C++ driver (same binary, same settings, N repetitions)
R="host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet"
for mode in spin block; do
for lanes in ${LANES:-1 4}; do
  line=""
  for rep in 1 2 3 4 5; do
    if [ $mode = block ]; then export SPC_BLOCKING_SYNC=1; else unset SPC_BLOCKING_SYNC; fi
    v=$($R --frames ${FRAMES:-192} --lanes $lanes 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f'%d['ms_per_frame'])")
    line="$line $v"
  done
  echo "$mode lanes $lanes ms/frame:$line"
done
done
