for d in 0 1 2 4 3 7; do echo -n "DBG=$d  "; W2L_CTC_DBG=$d python tools/microbench_ctc.py 8 1000 100 29; done
