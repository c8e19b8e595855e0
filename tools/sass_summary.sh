#!/bin/bash
# Blackwell-native instruction evidence: counts of tcgen05 / TMA / TMEM SASS mnemonics in the shipped library.
# usage: bash tools/sass_summary.sh > profiles/sass_summary.txt   (no GPU needed)
SO=wavtokenizer_b200/csrc/libwavtok_b200.so
echo "# cuobjdump -sass $SO | grep -o '<mnemonic>' | sort | uniq -c      ($(date -u +%F), $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA load; .MULTICAST = multicast::cluster,"
echo "# .2CTA = cta_group::2 completion on the peer's mbarrier), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMAPF = TMA L2 prefetch,"
echo "# STTM = tcgen05.st (A operands written to tensor memory), UTMASTG = bulk tensor store, SYNCS = mbarrier ops,"
echo "# UTCATOMSWS = tcgen05.alloc/dealloc, ELECT = elect.sync; UTCHMMA with a tmem[] A operand = A-from-TMEM form"
cuobjdump -sass $SO 2>/dev/null | grep -oE "UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMAPF[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|LDTM[.A-Z0-9_x]*|UTCATOMSWS[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|ELECT[.A-Z0-9_]*|UCGABAR_[A-Z]*|STTM[.A-Z0-9_x]*|UTMASTG[.A-Z0-9_]*|STG\.E\.ENL2\.256|MUFU\.EX2" | sort | uniq -c | sort -k2
echo "# kernels in the library (cuobjdump -sass | grep 'Function :' | c++filt):"
cuobjdump -sass $SO 2>/dev/null | grep -oE "Function : [A-Za-z0-9_]+" | sed 's/Function : //' | c++filt | sed -e 's/(anonymous namespace):://g' -e 's/^void //' -e 's/(.*//' | sort | uniq -c
echo "# A-from-TMEM MMAs (first operand tmem[..] instead of a shared-memory descriptor): $(cuobjdump -sass $SO 2>/dev/null | grep UTCHMMA | grep -vc 'gdesc\[UR[0-9]*\], gdesc') of $(cuobjdump -sass $SO 2>/dev/null | grep -c UTCHMMA) UTCHMMA"
