// Self-test of the MPI shim: run with  bin/mpirun -np N test/selftest  (N = 1, 2, 3, 5 ...).
#include <mpi.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static int rank_, size_, fails = 0;
#define CHECK(c)                                                                 \
  do {                                                                           \
    if (!(c)) {                                                                  \
      std::printf("[rank %d] FAIL %s:%d %s\n", rank_, __FILE__, __LINE__, #c);   \
      ++fails;                                                                   \
    }                                                                            \
  } while (0)

int main(int argc, char** argv) {
  int prov = 0;
  MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &prov);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank_);
  MPI_Comm_size(MPI_COMM_WORLD, &size_);
  const int P = size_, me = rank_;

  // allreduce, several sizes (incl. multi-chunk) and types
  for (long n : {1L, 3L, 1000L, 1L << 20, 3L << 20}) {
    std::vector<double> a(n), b(n);
    for (long i = 0; i < n; ++i) a[i] = (me + 1) * 0.5 + (i % 7);
    CHECK(MPI_Allreduce(a.data(), b.data(), (int)n, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD) == MPI_SUCCESS);
    for (long i = 0; i < n; i += (n > 100 ? 997 : 1)) CHECK(std::fabs(b[i] - (0.25 * P * (P + 1) + P * (i % 7))) < 1e-9);
    std::vector<float> f(n, (float)(me + 1)), g(n);
    MPI_Allreduce(f.data(), g.data(), (int)n, MPI_FLOAT, MPI_MAX, MPI_COMM_WORLD);
    CHECK(g[0] == (float)P && g[n - 1] == (float)P);
    std::vector<long> l(n, 1L << me), m(n);
    MPI_Allreduce(l.data(), m.data(), (int)n, MPI_LONG, MPI_BOR, MPI_COMM_WORLD);
    CHECK(m[n / 2] == (1L << P) - 1);
  }
  // in-place reduce to a root + bcast
  for (int root = 0; root < P; ++root) {
    std::vector<int> v(100, me + 1);
    if (me == root) MPI_Reduce(MPI_IN_PLACE, v.data(), 100, MPI_INT, MPI_SUM, root, MPI_COMM_WORLD);
    else MPI_Reduce(v.data(), nullptr, 100, MPI_INT, MPI_SUM, root, MPI_COMM_WORLD);
    if (me == root) CHECK(v[17] == P * (P + 1) / 2);
    std::vector<double> w(5000000, me == root ? 3.25 : 0.0);
    MPI_Bcast(w.data(), (int)w.size(), MPI_DOUBLE, root, MPI_COMM_WORLD);
    CHECK(w[0] == 3.25 && w[4999999] == 3.25);
  }
  // gatherv with vector+resized types along the middle axis of [before, axis, after]
  {
    const int before = 3, after = 4, mine = me + 1;
    std::vector<int> counts(P), displs(P);
    MPI_Allgather(&mine, 1, MPI_INT, counts.data(), 1, MPI_INT, MPI_COMM_WORLD);
    int tot = 0;
    for (int r = 0; r < P; ++r) {
      displs[r] = tot;
      tot += counts[r];
    }
    std::vector<float> in(before * mine * after);
    for (int b = 0; b < before; ++b)
      for (int i = 0; i < mine; ++i)
        for (int k = 0; k < after; ++k) in[(b * mine + i) * after + k] = 1000 * me + 100 * b + 10 * i + k;
    MPI_Datatype t1, st, t2, rt;
    MPI_Type_vector(before, after, after * mine, MPI_FLOAT, &t1);
    MPI_Type_create_resized(t1, 0, after * sizeof(float), &st);
    MPI_Type_commit(&st);
    MPI_Type_vector(before, after, after * tot, MPI_FLOAT, &t2);
    MPI_Type_create_resized(t2, 0, after * sizeof(float), &rt);
    MPI_Type_commit(&rt);
    std::vector<float> out(before * tot * after, -1.f);
    CHECK(MPI_Allgatherv(in.data(), mine, st, out.data(), counts.data(), displs.data(), rt, MPI_COMM_WORLD) == MPI_SUCCESS);
    for (int r = 0; r < P; ++r)
      for (int b = 0; b < before; ++b)
        for (int i = 0; i < counts[r]; ++i)
          for (int k = 0; k < after; ++k) CHECK(out[(b * tot + displs[r] + i) * after + k] == 1000 * r + 100 * b + 10 * i + k);
    std::vector<float> out2(before * tot * after, -1.f);
    MPI_Gatherv(in.data(), mine, st, out2.data(), counts.data(), displs.data(), rt, P - 1, MPI_COMM_WORLD);
    if (me == P - 1) CHECK(out2 == out);
    // scatter it back
    std::vector<float> back(before * mine * after, -2.f);
    MPI_Scatterv(out.data(), counts.data(), displs.data(), rt, back.data(), mine, st, 0, MPI_COMM_WORLD);
    CHECK(back == in);
    MPI_Type_free(&t1);
    MPI_Type_free(&t2);
    MPI_Type_free(&st);
    MPI_Type_free(&rt);
  }
  // 80 MB ring: Isend right, Irecv left, wait both; then tags out of order
  {
    const int n = 10000000;
    std::vector<double> s(n, me + 0.5), r(n, -1);
    MPI_Request rs, rr;
    MPI_Isend(s.data(), n, MPI_DOUBLE, (me + 1) % P, 7, MPI_COMM_WORLD, &rs);
    MPI_Irecv(r.data(), n, MPI_DOUBLE, (me + P - 1) % P, 7, MPI_COMM_WORLD, &rr);
    MPI_Status st;
    MPI_Wait(&rr, &st);
    MPI_Wait(&rs, MPI_STATUS_IGNORE);
    CHECK(r[0] == (me + P - 1) % P + 0.5 && r[n - 1] == r[0]);
    int cnt = 0;
    MPI_Get_count(&st, MPI_DOUBLE, &cnt);
    CHECK(cnt == n && st.MPI_TAG == 7);
    int a = 100 + me, b = 200 + me, ra = 0, rb = 0;
    MPI_Request q[4];
    MPI_Isend(&a, 1, MPI_INT, (me + 1) % P, 1, MPI_COMM_WORLD, &q[0]);
    MPI_Isend(&b, 1, MPI_INT, (me + 1) % P, 2, MPI_COMM_WORLD, &q[1]);
    MPI_Irecv(&rb, 1, MPI_INT, (me + P - 1) % P, 2, MPI_COMM_WORLD, &q[2]);  // tag 2 first: tag 1 becomes unexpected
    MPI_Wait(&q[2], MPI_STATUS_IGNORE);
    MPI_Irecv(&ra, 1, MPI_INT, (me + P - 1) % P, 1, MPI_COMM_WORLD, &q[3]);
    MPI_Waitall(4, q, MPI_STATUSES_IGNORE);
    CHECK(ra == 100 + (me + P - 1) % P && rb == 200 + (me + P - 1) % P);
  }
  int total = 0;
  MPI_Allreduce(&fails, &total, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  if (me == 0) std::printf("mpishim selftest np=%d: %s\n", P, total ? "FAILED" : "ok");
  MPI_Finalize();
  return total ? 1 : 0;
}
