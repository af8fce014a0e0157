export LD_LIBRARY_PATH=$PWD/opensot_b200/lib:$PWD/oracle/_ref
./tests/cpp/_bin/test_dropin nhqp C29 64 0.2 31 2>&1 | grep -v "^\[nHQP\]\|INFO\|^$" | tail -40
./tests/cpp/_bin/test_dropin nhqp C29 8 0.2 31 2>&1 | grep -v "^\[nHQP\]\|INFO\|^$" | tail -10 | cut -c1-300
