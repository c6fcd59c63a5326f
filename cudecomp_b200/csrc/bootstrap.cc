// TCP-mesh bootstrap. See bootstrap.h.
#include "bootstrap.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>

namespace cdb {

namespace {

struct World {
  bool initialized = false;
  int rank = 0;
  int size = 1;
  std::vector<int> fds; // socket to each world rank, -1 for self
  CommPtr world;
  CommPtr self;
  int next_comm_id = 16;
};

World g_world;

[[noreturn]] void fail(const std::string& what) {
  throw BootstrapError("bootstrap: " + what + (errno ? std::string(" (") + std::strerror(errno) + ")" : ""));
}

int envInt(const char* name, int dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  return std::atoi(v);
}

void setNoDelay(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}

void sendAll(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n > 0) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) {
      if (errno == EINTR) continue;
      fail("send failed");
    }
    p += k;
    n -= static_cast<size_t>(k);
  }
}

void recvAll(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n > 0) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k < 0) {
      if (errno == EINTR) continue;
      fail("recv failed");
    }
    if (k == 0) {
      errno = 0;
      fail("peer closed the connection (a rank exited early?)");
    }
    p += k;
    n -= static_cast<size_t>(k);
  }
}

// First words of a freshly accepted connection. Anything may dial the rendezvous port (a port scanner, a health
// check, a rank of another job): a connection that does not deliver `n` bytes within a few seconds, or closes early,
// is not a rank of this job. Returns false instead of failing the job; the caller drops the connection and keeps
// accepting.
bool recvGreeting(int fd, void* buf, size_t n) {
  timeval tv{5, 0};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  char* p = static_cast<char*>(buf);
  bool ok = true;
  while (n > 0) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) { // timeout (EAGAIN), reset or orderly close before the greeting was complete
      ok = false;
      break;
    }
    p += k;
    n -= static_cast<size_t>(k);
  }
  timeval off{0, 0};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &off, sizeof(off));
  return ok;
}

struct MsgHeader {
  uint32_t comm_id;
  uint32_t seq;
  uint64_t bytes;
};

void sendMsg(int peer, const Comm& c, const void* buf, size_t bytes) {
  MsgHeader h{static_cast<uint32_t>(c.id), c.seq, bytes};
  sendAll(g_world.fds[peer], &h, sizeof(h));
  if (bytes) sendAll(g_world.fds[peer], buf, bytes);
}

void recvMsg(int peer, const Comm& c, void* buf, size_t bytes) {
  MsgHeader h;
  recvAll(g_world.fds[peer], &h, sizeof(h));
  if (h.comm_id != static_cast<uint32_t>(c.id) || h.seq != c.seq || h.bytes != bytes) {
    errno = 0;
    fail("collective mismatch between ranks (comm " + std::to_string(c.id) + " seq " + std::to_string(c.seq) +
         " expected " + std::to_string(bytes) + " bytes, peer " + std::to_string(peer) + " sent comm " +
         std::to_string(h.comm_id) + " seq " + std::to_string(h.seq) + " bytes " + std::to_string(h.bytes) +
         "); collectives must be called in the same order on every rank");
  }
  if (bytes) recvAll(g_world.fds[peer], buf, bytes);
}

// Listens on `ip_be` only (network byte order; the library is single-node, there is no reason to accept connections on
// every interface). 0 or an address that is not local to this host (NAT, container port mapping) falls back to all
// interfaces.
int listenOn(uint32_t ip_be, int port, int* bound_port) {
  int fd = ::socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) fail("socket");
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<uint16_t>(port));
  sa.sin_addr.s_addr = ip_be ? ip_be : htonl(INADDR_ANY);
  if (::bind(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) < 0) {
    if (!ip_be) fail("bind to port " + std::to_string(port));
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    if (::bind(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) < 0) fail("bind to port " + std::to_string(port));
  }
  if (::listen(fd, 512) < 0) fail("listen");
  socklen_t len = sizeof(sa);
  getsockname(fd, reinterpret_cast<sockaddr*>(&sa), &len);
  if (bound_port) *bound_port = ntohs(sa.sin_port);
  return fd;
}

uint32_t resolve(const std::string& host) {
  addrinfo hints{};
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  addrinfo* res = nullptr;
  if (getaddrinfo(host.c_str(), nullptr, &hints, &res) != 0 || !res) {
    errno = 0;
    fail("cannot resolve " + host);
  }
  uint32_t ip = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr.s_addr;
  freeaddrinfo(res);
  return ip;
}

// Every rendezvous and mesh connection opens with this word: ranks of another job (or anything else that finds the
// port) are turned away instead of being taken for a rank. CUDECOMP_B200_BOOTSTRAP_TOKEN supplies a secret; without one
// the word is derived from what the launcher gave every rank of THIS job (not a secret, a mix-up guard).
uint64_t jobToken(const std::string& addr, int port, int size) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const std::string& s) {
    for (unsigned char ch : s) h = (h ^ ch) * 1099511628211ull;
    h = (h ^ 0xffu) * 1099511628211ull;
  };
  if (const char* t = std::getenv("CUDECOMP_B200_BOOTSTRAP_TOKEN")) mix(t);
  if (const char* t = std::getenv("TORCHELASTIC_RUN_ID")) mix(t);
  mix(addr);
  mix(std::to_string(port));
  mix(std::to_string(size));
  return h;
}

int connectTo(uint32_t ip_be, int port, double timeout_s) {
  auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) fail("socket");
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_port = htons(static_cast<uint16_t>(port));
    sa.sin_addr.s_addr = ip_be;
    if (::connect(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) == 0) {
      setNoDelay(fd);
      return fd;
    }
    ::close(fd);
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (el > timeout_s) fail("connect to rendezvous port " + std::to_string(port) + " timed out");
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
}

int acceptOne(int lfd) {
  for (;;) {
    int fd = ::accept(lfd, nullptr, nullptr);
    if (fd >= 0) {
      setNoDelay(fd);
      return fd;
    }
    if (errno == EINTR) continue;
    fail("accept");
  }
}

struct PeerAddr {
  uint32_t ip_be;
  int32_t port;
};

} // namespace

bool worldInitialized() { return g_world.initialized; }
int worldRank() { return g_world.rank; }
int worldSize() { return g_world.size; }
CommPtr worldComm() { return g_world.world; }
CommPtr selfComm() { return g_world.self; }

void worldInit() {
  if (g_world.initialized) return;
  const int rank = envInt("RANK", envInt("OMPI_COMM_WORLD_RANK", envInt("PMI_RANK", 0)));
  const int size = envInt("WORLD_SIZE", envInt("OMPI_COMM_WORLD_SIZE", envInt("PMI_SIZE", 1)));
  const char* addr_env = std::getenv("CUDECOMP_B200_BOOTSTRAP_ADDR");
  if (!addr_env) addr_env = std::getenv("MASTER_ADDR");
  int port = envInt("CUDECOMP_B200_BOOTSTRAP_PORT", 0);
  if (port == 0) port = envInt("MASTER_PORT", 29616) + 1;
  worldInitExplicit(rank, size, addr_env ? addr_env : "127.0.0.1", port);
}

void worldInitExplicit(int rank, int size, const std::string& addr, int port) {
  if (g_world.initialized) {
    if (g_world.rank != rank || g_world.size != size) {
      errno = 0;
      fail("the process mesh already exists with a different rank / size");
    }
    return;
  }
  World& w = g_world;
  w.rank = rank;
  w.size = size;
  if (w.size < 1 || w.rank < 0 || w.rank >= w.size) {
    errno = 0;
    fail("invalid rank / size (RANK, WORLD_SIZE)");
  }
  w.fds.assign(w.size, -1);

  if (w.size > 1) {
    double timeout_s = envInt("CUDECOMP_B200_BOOTSTRAP_TIMEOUT", 120);

    int my_port = 0;
    int lfd = -1;
    std::vector<PeerAddr> table(w.size);
    const uint64_t token = jobToken(addr, port, w.size);
    struct Hello {
      uint64_t token;
      int32_t rank;
      int32_t port;
    };
    if (w.rank == 0) {
      lfd = listenOn(resolve(addr), port, nullptr);
      // every other rank dials in and reports the port of its own listening socket
      for (int k = 1; k < w.size; ++k) {
        int fd = acceptOne(lfd);
        Hello hello{};
        if (!recvGreeting(fd, &hello, sizeof(hello)) || hello.token != token) { // not a rank of this job
          ::close(fd);
          --k;
          continue;
        }
        int r = hello.rank;
        if (r <= 0 || r >= w.size || w.fds[r] != -1) {
          errno = 0;
          fail("unexpected rank " + std::to_string(r) + " at rendezvous");
        }
        w.fds[r] = fd;
        sockaddr_in pa{};
        socklen_t len = sizeof(pa);
        getpeername(fd, reinterpret_cast<sockaddr*>(&pa), &len);
        table[r] = PeerAddr{pa.sin_addr.s_addr, hello.port};
      }
      table[0] = PeerAddr{0, port};
      for (int r = 1; r < w.size; ++r) sendAll(w.fds[r], table.data(), sizeof(PeerAddr) * w.size);
    } else {
      int fd = connectTo(resolve(addr), port, timeout_s);
      // my mesh listener: on the interface that reaches rank 0
      sockaddr_in la{};
      socklen_t llen = sizeof(la);
      getsockname(fd, reinterpret_cast<sockaddr*>(&la), &llen);
      lfd = listenOn(la.sin_addr.s_addr, 0, &my_port);
      Hello hello{token, w.rank, my_port};
      sendAll(fd, &hello, sizeof(hello));
      w.fds[0] = fd;
      recvAll(fd, table.data(), sizeof(PeerAddr) * w.size);
      // mesh among the non-zero ranks: the higher rank dials the lower one
      for (int r = 1; r < w.rank; ++r) {
        int pfd = connectTo(table[r].ip_be, table[r].port, timeout_s);
        Hello me{token, w.rank, 0};
        sendAll(pfd, &me, sizeof(me));
        w.fds[r] = pfd;
      }
      for (int k = w.rank + 1; k < w.size; ++k) {
        int pfd = acceptOne(lfd);
        Hello other{};
        if (!recvGreeting(pfd, &other, sizeof(other)) || other.token != token) {
          ::close(pfd);
          --k;
          continue;
        }
        const int32_t who = other.rank;
        if (who <= w.rank || who >= w.size || w.fds[who] != -1) {
          errno = 0;
          fail("unexpected rank in mesh setup");
        }
        w.fds[who] = pfd;
      }
    }
    ::close(lfd);
  }

  w.world = std::make_shared<Comm>();
  w.world->id = 1;
  w.world->members.resize(w.size);
  for (int i = 0; i < w.size; ++i) w.world->members[i] = i;
  w.world->me = w.rank;
  w.self = std::make_shared<Comm>();
  w.self->id = 2;
  w.self->members = {w.rank};
  w.self->me = 0;
  w.initialized = true;
  barrier(*w.world);
}

int pickFreePort() {
  int port = 0;
  int fd = listenOn(htonl(INADDR_LOOPBACK), 0, &port);
  ::close(fd);
  return port;
}

void worldFinalize() {
  if (!g_world.initialized) return;
  try {
    barrier(*g_world.world);
  } catch (...) {}
  for (int& fd : g_world.fds) {
    if (fd >= 0) ::close(fd);
    fd = -1;
  }
  g_world.world.reset();
  g_world.self.reset();
  g_world.initialized = false;
}

// Star through the first member: contributions in, full result out.
void allgather(Comm& c, const void* in, size_t bytes, void* out) {
  const int n = c.size();
  char* o = static_cast<char*>(out);
  if (n == 1) {
    if (out != in) std::memcpy(o, in, bytes);
    c.seq++;
    return;
  }
  if (c.me == 0) {
    std::memmove(o, in, bytes);
    for (int i = 1; i < n; ++i) recvMsg(c.members[i], c, o + i * bytes, bytes);
    for (int i = 1; i < n; ++i) sendMsg(c.members[i], c, o, bytes * n);
  } else {
    // `in` may alias a slot of `out`; send first
    sendMsg(c.members[0], c, in, bytes);
    recvMsg(c.members[0], c, o, bytes * n);
  }
  c.seq++;
}

void bcast(Comm& c, void* buf, size_t bytes, int root) {
  const int n = c.size();
  if (n > 1) {
    if (c.me == root) {
      for (int i = 0; i < n; ++i)
        if (i != root) sendMsg(c.members[i], c, buf, bytes);
    } else {
      recvMsg(c.members[root], c, buf, bytes);
    }
  }
  c.seq++;
}

void barrier(Comm& c) {
  char x = 0;
  std::vector<char> all(c.size());
  allgather(c, &x, 1, all.data());
}

CommPtr split(Comm& c, int color, int key) {
  struct Entry {
    int32_t color, key, idx;
  };
  std::vector<Entry> all(c.size());
  Entry mine{color, key, c.me};
  allgather(c, &mine, sizeof(Entry), all.data());
  // children of the same parent get distinct ids that are identical on all members
  uint32_t ordinal = c.nsplits++;
  if (color < 0) return nullptr;
  std::vector<Entry> grp;
  for (auto& e : all)
    if (e.color == color) grp.push_back(e);
  std::stable_sort(grp.begin(), grp.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
  auto out = std::make_shared<Comm>();
  // id: hash of (parent id, ordinal, color) -- equal across the members of the new group
  uint64_t h = 1469598103934665603ull;
  for (uint64_t v : {static_cast<uint64_t>(c.id), static_cast<uint64_t>(ordinal), static_cast<uint64_t>(color)}) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 1099511628211ull;
  }
  out->id = static_cast<int>((h & 0x7fffffff) | 0x100);
  for (size_t i = 0; i < grp.size(); ++i) {
    out->members.push_back(c.members[grp[i].idx]);
    if (grp[i].idx == c.me) out->me = static_cast<int>(i);
  }
  return out;
}

CommPtr dup(Comm& c) { return split(c, 0, c.me); }

namespace {
template <typename T> void reduceInto(T* acc, const T* v, int n, ReduceOp op) {
  for (int i = 0; i < n; ++i) {
    switch (op) {
    case ReduceOp::SUM: acc[i] += v[i]; break;
    case ReduceOp::PROD: acc[i] *= v[i]; break;
    case ReduceOp::MAX: acc[i] = std::max(acc[i], v[i]); break;
    case ReduceOp::MIN: acc[i] = std::min(acc[i], v[i]); break;
    case ReduceOp::LOR: acc[i] = (acc[i] != T(0) || v[i] != T(0)) ? T(1) : T(0); break;
    case ReduceOp::LAND: acc[i] = (acc[i] != T(0) && v[i] != T(0)) ? T(1) : T(0); break;
    case ReduceOp::BOR: acc[i] = static_cast<T>(static_cast<int64_t>(acc[i]) | static_cast<int64_t>(v[i])); break;
    }
  }
}
template <typename T> void allreduceT(Comm& c, T* v, int n, ReduceOp op) {
  std::vector<T> all(static_cast<size_t>(n) * c.size());
  allgather(c, v, sizeof(T) * n, all.data());
  // every rank reduces in member order, so the result is bitwise identical everywhere
  for (int i = 0; i < n; ++i) v[i] = all[i];
  for (int r = 1; r < c.size(); ++r) reduceInto(v, all.data() + static_cast<size_t>(r) * n, n, op);
}
} // namespace

void allreduceF64(Comm& c, double* v, int n, ReduceOp op) { allreduceT(c, v, n, op); }
void allreduceI64(Comm& c, int64_t* v, int n, ReduceOp op) { allreduceT(c, v, n, op); }

uint64_t sharedToken(Comm& c) {
  uint64_t tok = 0;
  if (c.me == 0) {
    std::random_device rd;
    tok = (static_cast<uint64_t>(rd()) << 32) ^ rd() ^ (static_cast<uint64_t>(::getpid()) << 16);
  }
  bcast(c, &tok, sizeof(tok), 0);
  return tok;
}

} // namespace cdb
