for f in scripts/build/libmsi_mb*.so; do MSI_B200_LIB=$PWD/$f timeout 100 python scripts/time_stages.py 2>&1 | tail -1; done
