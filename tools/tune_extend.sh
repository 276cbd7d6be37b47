for t in 24,12,6,4 24,16,8,6 24,20,10,8 24,24,12,12 20,16,8,6 28,16,8,6 24,16,4,6 24,8,4,2 24,32,16,32; do echo "tune $t"; PTC_EXTEND_TUNE=$t python tools/trav_stats.py 1 Atrium; done
