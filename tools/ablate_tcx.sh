for v in 0 15 8 2; do echo -n "ablate=$v: "; SIFNN_TC_ABLATE=$v python tools/profile_ops.py --time --tc --only 16x16x256 2>&1 | grep "16 @256" | sed 's/.*fwd_tc/fwd_tc/' | cut -c 1-120; done
