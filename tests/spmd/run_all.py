"""SPMD test runner: every rank runs every test of tests/spmd/spmd_*.py in the
same order (the reference runs ``mpirun -np N nose2`` the same way,
.github/workflows/test.yml:64-84).  Exit code != 0 if any test failed here."""
import os
import sys
import unittest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main() -> int:
    import mpi4torch_b200 as m4t

    comm = m4t.COMM_WORLD
    pattern = sys.argv[1] if len(sys.argv) > 1 else "spmd_*.py"
    suite = unittest.defaultTestLoader.discover(HERE, pattern=pattern)
    stream = sys.stderr if comm.rank == 0 else open(os.devnull, "w")
    result = unittest.TextTestRunner(stream=stream, verbosity=1 if comm.rank == 0 else 0).run(suite)
    ok = result.wasSuccessful()
    if not ok:
        for test, tb in result.failures + result.errors:
            sys.stderr.write(f"[rank {comm.rank}] FAILED {test}\n{tb}\n")
    sys.stderr.flush()
    if comm.rank == 0:
        print(f"SPMD suite np={comm.size} device={os.environ.get('M4T_TEST_DEVICE', 'cpu')}: "
              f"ran {result.testsRun} tests, ok={ok}", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
