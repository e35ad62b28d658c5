// ref_mpi_threads.cpp -- the MPI subset the reference calls, with the ranks as threads of this process
// (TEST INFRASTRUCTURE ONLY; see oracle/ref_stub/mpi.h).  Lets the reference's own FiniteVolumeGrid2D::partition,
// initCommBuffers, sendMessages and IndexMap run on several "ranks" inside oracle/_ref/libphase_ref_fv.so.
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <mpi.h>

namespace {
thread_local int t_rank = 0;
int g_size = 1;
std::mutex g_m;
std::condition_variable g_cv;

int type_size(MPI_Datatype t) {
  if (t >= MPI_DERIVED_BASE_) return t - MPI_DERIVED_BASE_;
  return t == MPI_BYTE ? 1 : t == MPI_INT ? 4 : 8;
}

// ---- collectives: everybody deposits a pointer, meets, reads, meets again
struct Board {
  std::vector<const void *> ptr;
  std::vector<long> len;
  int arrived = 0;
  long gen = 0;
} g_board;

void meet(std::unique_lock<std::mutex> &lk) {
  if (++g_board.arrived == g_size) {
    g_board.arrived = 0;
    g_board.gen++;
    g_cv.notify_all();
  } else {
    const long g = g_board.gen;
    g_cv.wait(lk, [&] { return g_board.gen != g; });
  }
}
// deposit (p, n) and wait for everybody; on return the board is readable until the closing meet()
void deposit(std::unique_lock<std::mutex> &lk, const void *p, long n) {
  g_board.ptr[t_rank] = p;
  g_board.len[t_rank] = n;
  meet(lk);
}

// ---- point to point
struct Msg { int tag; std::vector<char> data; bool sync; bool *consumed; };
struct Posted { void *buf; long cap; int src, tag; bool done; long got; };
std::vector<std::vector<std::deque<Msg>>> g_box;         // [dst][src]
std::vector<std::deque<Posted *>> g_posted;                // [dst]
thread_local std::vector<Posted *> t_requests;             // index = request id - 1

bool tag_ok(int want, int have) { return want == MPI_ANY_TAG || have == MPI_ANY_TAG || want == have; }

void send(const void *buf, long bytes, int dest, int tag, bool sync) {
  std::unique_lock<std::mutex> lk(g_m);
  for (auto it = g_posted[dest].begin(); it != g_posted[dest].end(); ++it) {
    Posted *p = *it;
    if (!p->done && p->src == t_rank && tag_ok(p->tag, tag)) {   // the matching receive is posted: deliver
      memcpy(p->buf, buf, (size_t)std::min(bytes, p->cap));
      p->got = bytes;
      p->done = true;
      g_posted[dest].erase(it);
      g_cv.notify_all();
      return;
    }
  }
  bool consumed = false;
  Msg m;
  m.tag = tag; m.data.assign((const char *)buf, (const char *)buf + bytes); m.sync = sync; m.consumed = sync ? &consumed : nullptr;
  g_box[dest][t_rank].push_back(std::move(m));
  g_cv.notify_all();
  if (sync) g_cv.wait(lk, [&] { return consumed; });
}
// first matching message from `source` to me, or nullptr
Msg *find_msg(int source, int tag) {
  for (Msg &m : g_box[t_rank][source])
    if (tag_ok(tag, m.tag)) return &m;
  return nullptr;
}
void take(int source, Msg *m, void *buf, long cap, long *got) {
  memcpy(buf, m->data.data(), (size_t)std::min<long>((long)m->data.size(), cap));
  if (got) *got = (long)m->data.size();
  if (m->consumed) *m->consumed = true;
  auto &q = g_box[t_rank][source];
  for (auto it = q.begin(); it != q.end(); ++it)
    if (&*it == m) { q.erase(it); break; }
  g_cv.notify_all();
}
}  // namespace

int MPI_Init(int *, char ***) { return 0; }
int MPI_Finalize() { return 0; }
int MPI_Comm_rank(MPI_Comm, int *r) { *r = t_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int *n) { *n = g_size; return 0; }
int MPI_Type_vector(int count, int blocklen, int, MPI_Datatype t, MPI_Datatype *out) {
  *out = MPI_DERIVED_BASE_ + count * blocklen * type_size(t);
  return 0;
}
int MPI_Type_commit(MPI_Datatype *) { return 0; }

int MPI_Barrier(MPI_Comm) {
  if (g_size == 1) return 0;
  std::unique_lock<std::mutex> lk(g_m);
  meet(lk);
  return 0;
}
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm) {
  if (g_size == 1) return 0;
  std::unique_lock<std::mutex> lk(g_m);
  deposit(lk, buf, (long)n * type_size(t));
  if (t_rank != root) memcpy(buf, g_board.ptr[root], (size_t)n * type_size(t));
  meet(lk);
  return 0;
}
int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm) {
  const size_t bytes = (size_t)n * type_size(t);
  if (g_size == 1) { if (in != out) memcpy(out, in, bytes); return 0; }
  std::vector<char> tmp(bytes);
  std::unique_lock<std::mutex> lk(g_m);
  deposit(lk, in, (long)bytes);
  auto reduce = [&](auto *res, auto zero) {
    typedef decltype(zero) V;
    for (int i = 0; i < n; ++i) {
      V acc = ((const V *)g_board.ptr[0])[i];
      for (int q = 1; q < g_size; ++q) {
        const V v = ((const V *)g_board.ptr[q])[i];
        acc = op == MPI_SUM ? acc + v : op == MPI_MIN ? (v < acc ? v : acc) : (v > acc ? v : acc);
      }
      res[i] = acc;
    }
  };
  if (t == MPI_DOUBLE) reduce((double *)tmp.data(), 0.0);
  else if (t == MPI_INT) reduce((int *)tmp.data(), 0);
  else if (t == MPI_LONG) reduce((long *)tmp.data(), 0L);
  else if (t == MPI_UNSIGNED_LONG) reduce((unsigned long *)tmp.data(), 0UL);
  else abort();
  meet(lk);
  memcpy(out, tmp.data(), bytes);
  return 0;
}
int MPI_Allgatherv(const void *in, int n, MPI_Datatype t, void *out, const int *counts, const int *displs, MPI_Datatype tr, MPI_Comm) {
  const int ts = type_size(tr);
  if (g_size == 1) { memcpy((char *)out + (size_t)(displs ? displs[0] : 0) * ts, in, (size_t)n * type_size(t)); return 0; }
  std::unique_lock<std::mutex> lk(g_m);
  deposit(lk, in, (long)n * type_size(t));
  for (int q = 0; q < g_size; ++q) memcpy((char *)out + (size_t)displs[q] * ts, g_board.ptr[q], (size_t)g_board.len[q]);
  (void)counts;
  meet(lk);
  return 0;
}
int MPI_Allgather(const void *in, int n, MPI_Datatype t, void *out, int, MPI_Datatype, MPI_Comm) {
  const size_t bytes = (size_t)n * type_size(t);
  if (g_size == 1) { memcpy(out, in, bytes); return 0; }
  std::unique_lock<std::mutex> lk(g_m);
  deposit(lk, in, (long)bytes);
  for (int q = 0; q < g_size; ++q) memcpy((char *)out + q * bytes, g_board.ptr[q], bytes);
  meet(lk);
  return 0;
}
int MPI_Gather(const void *in, int n, MPI_Datatype t, void *out, int, MPI_Datatype, int root, MPI_Comm) {
  const size_t bytes = (size_t)n * type_size(t);
  if (g_size == 1) { memcpy(out, in, bytes); return 0; }
  std::unique_lock<std::mutex> lk(g_m);
  deposit(lk, in, (long)bytes);
  if (t_rank == root) for (int q = 0; q < g_size; ++q) memcpy((char *)out + q * bytes, g_board.ptr[q], bytes);
  meet(lk);
  return 0;
}
int MPI_Gatherv(const void *in, int n, MPI_Datatype t, void *out, const int *, const int *displs, MPI_Datatype tr, int root, MPI_Comm) {
  const int ts = type_size(tr);
  if (g_size == 1) { memcpy((char *)out + (size_t)(displs ? displs[0] : 0) * ts, in, (size_t)n * type_size(t)); return 0; }
  std::unique_lock<std::mutex> lk(g_m);
  deposit(lk, in, (long)n * type_size(t));
  if (t_rank == root) for (int q = 0; q < g_size; ++q) memcpy((char *)out + (size_t)displs[q] * ts, g_board.ptr[q], (size_t)g_board.len[q]);
  meet(lk);
  return 0;
}

int MPI_Ssend(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm) {
  if (g_size == 1) abort();
  send(buf, (long)n * type_size(t), dest, tag, true);
  return 0;
}
int MPI_Isend(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request *req) {
  if (g_size == 1) abort();
  send(buf, (long)n * type_size(t), dest, tag, false);   // buffered: complete on return
  *req = 0;
  return 0;
}
int MPI_Irecv(void *buf, int n, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Request *req) {
  if (g_size == 1) abort();
  std::unique_lock<std::mutex> lk(g_m);
  Posted *p = new Posted{buf, (long)n * type_size(t), source, tag, false, 0};
  if (Msg *m = find_msg(source, tag)) {
    take(source, m, buf, p->cap, &p->got);
    p->done = true;
  } else {
    g_posted[t_rank].push_back(p);
  }
  t_requests.push_back(p);
  *req = (int)t_requests.size();
  return 0;
}
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *) {
  if (n == 0) return 0;
  std::unique_lock<std::mutex> lk(g_m);
  for (int i = 0; i < n; ++i) {
    if (reqs[i] <= 0) continue;
    Posted *p = t_requests[reqs[i] - 1];
    g_cv.wait(lk, [&] { return p->done; });
  }
  bool all = true;
  for (Posted *p : t_requests) all = all && p->done;
  if (all) { for (Posted *p : t_requests) delete p; t_requests.clear(); }
  return 0;
}
int MPI_Probe(int source, int tag, MPI_Comm, MPI_Status *st) {
  if (g_size == 1) abort();
  std::unique_lock<std::mutex> lk(g_m);
  Msg *m = nullptr;
  g_cv.wait(lk, [&] { return (m = find_msg(source, tag)) != nullptr; });
  if (st) { st->MPI_SOURCE = source; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->bytes_ = (int)m->data.size(); }
  return 0;
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) {
  *count = st->bytes_ / type_size(t);
  return 0;
}
int MPI_Recv(void *buf, int n, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Status *st) {
  if (g_size == 1) abort();
  std::unique_lock<std::mutex> lk(g_m);
  Msg *m = nullptr;
  g_cv.wait(lk, [&] { return (m = find_msg(source, tag)) != nullptr; });
  long got = 0;
  const int mtag = m->tag;
  take(source, m, buf, (long)n * type_size(t), &got);
  if (st) { st->MPI_SOURCE = source; st->MPI_TAG = mtag; st->MPI_ERROR = 0; st->bytes_ = (int)got; }
  return 0;
}

void phase_mpi_run(int nRanks, void (*f)(int, void *), void *user) {
  g_size = nRanks;
  g_board.ptr.assign(nRanks, nullptr);
  g_board.len.assign(nRanks, 0);
  g_board.arrived = 0;
  g_box.assign(nRanks, std::vector<std::deque<Msg>>(nRanks));
  g_posted.assign(nRanks, std::deque<Posted *>());
  std::vector<std::thread> th;
  for (int r = 0; r < nRanks; ++r)
    th.emplace_back([=] { t_rank = r; t_requests.clear(); f(r, user); });
  for (auto &t : th) t.join();
  g_size = 1;
  t_rank = 0;
}
