#!/bin/sh
# Regenerates the committed known-answer files from the UNMODIFIED reference (oracle/_ref/katdump,
# built by oracle/Makefile from /root/reference).  Inputs are seeded (splitmix64) inside katdump.
set -e
K=../../oracle/_ref/katdump
$K extend 1500 12345 | gzip -9 > extend.kat.gz   # ksw_extend2  (ksw.c:416)
$K global 1000 6789  | gzip -9 > global.kat.gz   # ksw_global2  (ksw.c:540)
$K local  400 424242 | gzip -9 > local.kat.gz    # ksw_align2   (ksw.c:379)
