for d in 0 1 2 3; do echo "=== dbg $d (1 = no MMAs, 2 = no TMEM->xs conversion)"; AR_X_DBG=$d python tools/_probe_lstm2.py; done
