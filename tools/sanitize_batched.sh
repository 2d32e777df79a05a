# compute-sanitizer over the batched-generation path of nraps_mc_run (several generations per launch)
mkdir -p /tmp/san
for tool in memcheck racecheck; do
  for tr in surface woodcock; do
    for deck in a c; do
      echo "== $tool $tr case_$deck (H=3000, 40 generations)"
      timeout 300 compute-sanitizer --tool $tool nraps_b200/lib/nraps tests/golden/decks/case_$deck.txt --histories 3000 --generations 40 --skip 2 \
        --tracking $tr --out /tmp/san --quiet 2>&1 | grep -E "k_fund|ERROR SUMMARY|Error|error" | head -5
    done
  done
done
