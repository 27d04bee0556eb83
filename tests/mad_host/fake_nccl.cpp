// fake_nccl.cpp -- TEST INFRASTRUCTURE.  The handful of NCCL entry points libmadgpu binds with dlopen (csrc/madgpu.cu, struct
// Nccl), for ranks that are THREADS of one process (tests/test_cpu_mad_host_slabs.py): sends are buffered in per-pair mailboxes,
// receives block until the matching message is there, all-reduce / broadcast meet at a generation barrier.  "Streams" are
// synchronous in the host build, so an operation has completed when the call returns.  Selected with MADGPU_NCCL_LIB.
// A receive or a collective that waits longer than FAKE_NCCL_TIMEOUT_S (default 120 s) reports who waits for whom and aborts.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

namespace
{
struct World {
  int nranks = 0, joined = 0, left = 0;
  std::mutex m;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<std::vector<char> > > mail;  // (src, dst) -> messages in order
  // collectives
  unsigned long coll_gen = 0;
  int coll_arrived = 0;
  std::vector<double> acc;
  std::vector<char> bcast;
};
struct Comm {
  World* w;
  int rank;
};
std::mutex g_m;
std::map<std::string, World*> g_worlds;

double timeout_s()
{
  const char* e = std::getenv("FAKE_NCCL_TIMEOUT_S");
  return e ? std::atof(e) : 120.0;
}
size_t dsize(int dtype) { return dtype == 8 ? 8 : dtype == 7 ? 4 : 0; }  // ncclFloat64 = 8, ncclFloat32 = 7

template <typename Pred>
void wait_or_die(World* w, std::unique_lock<std::mutex>& lk, Pred p, const char* what, int rank, int peer)
{
  if (!w->cv.wait_for(lk, std::chrono::duration<double>(timeout_s()), p)) {
    std::fprintf(stderr, "fake_nccl: rank %d timed out in %s (peer %d) -- a hung collective on a GPU\n", rank, what, peer);
    std::abort();
  }
}
}  // namespace

extern "C" {

struct ncclUniqueId { char internal[128]; };
typedef Comm* ncclComm_t;

int ncclGetUniqueId(ncclUniqueId* id)
{
  std::random_device rd;
  for (int i = 0; i < 128; ++i) id->internal[i] = (char)(rd() & 0xff);
  return 0;
}

int ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank)
{
  World* w;
  {
    std::lock_guard<std::mutex> g(g_m);
    World*& slot = g_worlds[std::string(id.internal, 128)];
    if (!slot) { slot = new World(); slot->nranks = nranks; }
    w = slot;
  }
  std::unique_lock<std::mutex> lk(w->m);
  ++w->joined;
  w->cv.notify_all();
  wait_or_die(w, lk, [&] { return w->joined >= w->nranks; }, "ncclCommInitRank", rank, -1);  // collective, like the real one
  *comm = new Comm{w, rank};
  return 0;
}

int ncclCommDestroy(ncclComm_t c)
{
  delete c;
  return 0;
}

int ncclGroupStart() { return 0; }
int ncclGroupEnd() { return 0; }
const char* ncclGetErrorString(int r) { return r == 0 ? "no error" : "fake NCCL error"; }

int ncclSend(const void* buf, size_t count, int dtype, int peer, ncclComm_t c, void*)
{
  const size_t bytes = count * dsize(dtype);
  if (!bytes && count) return 4;
  std::vector<char> msg((const char*)buf, (const char*)buf + bytes);
  std::lock_guard<std::mutex> g(c->w->m);
  c->w->mail[{c->rank, peer}].push_back(std::move(msg));
  c->w->cv.notify_all();
  return 0;
}

int ncclRecv(void* buf, size_t count, int dtype, int peer, ncclComm_t c, void*)
{
  const size_t bytes = count * dsize(dtype);
  World* w = c->w;
  std::unique_lock<std::mutex> lk(w->m);
  auto& q = w->mail[{peer, c->rank}];
  wait_or_die(w, lk, [&] { return !q.empty(); }, "ncclRecv", c->rank, peer);
  if (q.front().size() != bytes) {
    std::fprintf(stderr, "fake_nccl: rank %d receives %zu bytes from %d but %zu were sent\n", c->rank, bytes, peer, q.front().size());
    std::abort();
  }
  std::memcpy(buf, q.front().data(), bytes);
  q.pop_front();
  return 0;
}

// sum only; the ranks' contributions are added in ARRIVAL order, like a real all-reduce the result is the same on every rank
int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, ncclComm_t c, void*)
{
  if (op != 0 || !dsize(dtype)) return 4;
  World* w = c->w;
  std::unique_lock<std::mutex> lk(w->m);
  const unsigned long gen = w->coll_gen;
  if (w->coll_arrived == 0) w->acc.assign(count, 0.0);
  for (size_t i = 0; i < count; ++i) w->acc[i] += dtype == 8 ? ((const double*)send)[i] : (double)((const float*)send)[i];
  if (++w->coll_arrived == w->nranks) {
    w->coll_arrived = 0;
    ++w->coll_gen;
    w->cv.notify_all();
  } else {
    wait_or_die(w, lk, [&] { return w->coll_gen != gen; }, "ncclAllReduce", c->rank, -1);
  }
  // acc stays valid until the next collective's first arrival, which cannot happen before every rank has left this one:
  // count the leavers
  for (size_t i = 0; i < count; ++i) {
    if (dtype == 8) ((double*)recv)[i] = w->acc[i];
    else ((float*)recv)[i] = (float)w->acc[i];
  }
  if (++w->left == w->nranks) { w->left = 0; w->cv.notify_all(); }
  else wait_or_die(w, lk, [&] { return w->left == 0; }, "ncclAllReduce (exit)", c->rank, -1);
  return 0;
}

int ncclBroadcast(const void* send, void* recv, size_t count, int dtype, int root, ncclComm_t c, void*)
{
  const size_t bytes = count * dsize(dtype);
  World* w = c->w;
  std::unique_lock<std::mutex> lk(w->m);
  const unsigned long gen = w->coll_gen;
  if (c->rank == root) w->bcast.assign((const char*)send, (const char*)send + bytes);
  if (++w->coll_arrived == w->nranks) {
    w->coll_arrived = 0;
    ++w->coll_gen;
    w->cv.notify_all();
  } else {
    wait_or_die(w, lk, [&] { return w->coll_gen != gen; }, "ncclBroadcast", c->rank, root);
  }
  std::memcpy(recv, w->bcast.data(), bytes);
  if (++w->left == w->nranks) { w->left = 0; w->cv.notify_all(); }
  else wait_or_die(w, lk, [&] { return w->left == 0; }, "ncclBroadcast (exit)", c->rank, root);
  return 0;
}

}  // extern "C"
