for m in 0x08 0x24 0x88 0x92 0xAA; do echo "== mask $m"; DESTA_B200_LIB=tools/_bin/libdesta_polyb$m.so python tools/kbench.py attn 2>&1 | grep -E 'enc_unit'; done
