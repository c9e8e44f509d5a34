# development aid: time the decode launch with knock-out variants of the library (see profiles/r2_decode_ws_roles.md)
for v in "$@"; do echo "== $v"; SSV_B200_LIB=spoofsv_b200/build/variants/$v/libspoofsv_b200.so timeout 100 python tools/decode_time.py ${KO_SHAPES:-1 64 128 128:4:1:4} 2>&1 | grep -v Warn; done
