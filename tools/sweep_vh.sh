for cfg in C2 C1 C3; do
for n in 2 3 4 5 6 8 10 14; do for b in 2 4 6 8 12 16 24; do
echo -n "$cfg NCB=$n bands=$b: "; B200S_VH_VERBOSE=1 B200S_VH_NCB=$n B200S_VH_BANDS=$b python tools/time_bm.py $cfg 5 2>&1 | grep -E "plan|bm " | sort -u | sed 's/.*grid=/grid=/; s/.*: bm/bm/' | tr '\n' ' '; echo; done; done; done
