python -m pytest tests/test_gpu_overlap.py -x -q 2>&1 | tail -5
python scripts/overlap_probe.py 10 3 2>&1 | tail -8
MDB_KF_NSB=7 python scripts/overlap_probe.py 10 3 2>&1 | tail -8
MDB_KF_NSB=6 python scripts/overlap_probe.py 10 3 2>&1 | tail -8
MOLDY_B200_LIB=$PWD/moldy_b200/var/libmdb_mcw9.so MDB_KF_NSB=6 python scripts/overlap_probe.py 10 3 2>&1 | tail -8
