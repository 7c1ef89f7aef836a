// The processes of a slab-decomposed run: `kspaceFirstOrder-B200 --gpus N` forks N - 1 workers, ONE PROCESS PER GPU (rank r drives
// device r and owns the z-planes [r Nz/N, (r+1) Nz/N) of every grid).  The data plane between the GPUs is inside the engine library
// (all-to-all over NVLink, csrc/peer_dl.h); this class is only the host-side control plane: rank 0 owns the output and checkpoint
// files, the other ranks hand it their sensor rows, aggregate buffers and slabs through socket pairs created before the fork.
// Every rank executes the same host code in lock step, so a message needs no tag: the next thing rank 0 reads from rank r is the next
// thing rank r wrote.  A rank that dies closes its socket; its peers then fail with an error instead of waiting for ever.
//
// No counterpart in the reference, which is single-GPU (main.cpp:840-966; MatrixContainer holds whole arrays).
#pragma once
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kwhost {

class Team {
 public:
  int rank = 0, size = 1;
  bool multi() const { return size > 1; }
  bool root() const { return rank == 0; }

  // forks size - 1 workers; returns in every process with its rank set
  void spawn(int n) {
    size = n < 1 ? 1 : n;
    if (size == 1) return;
    std::vector<int> parentEnd(size, -1);
    for (int r = 1; r < size; ++r) {
      int sv[2];
      if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) throw std::runtime_error("Error: cannot create the socket pair of rank " + std::to_string(r) + ".");
      const pid_t pid = fork();
      if (pid < 0) throw std::runtime_error("Error: cannot fork the process of rank " + std::to_string(r) + ".");
      if (pid == 0) {  // worker: keeps only its own end
        for (int q = 1; q < r; ++q) close(parentEnd[q]);
        close(sv[0]);
        rank = r;
        mFd.assign(1, sv[1]);
        return;
      }
      close(sv[1]);
      parentEnd[r] = sv[0];
      mChildren.push_back(pid);
    }
    mFd = parentEnd;
  }

  // rank 0 <-> rank r (r > 0); workers only talk to rank 0
  void send(int r, const void* p, size_t n) const { io(fdOf(r), const_cast<void*>(p), n, true, r); }
  void recv(int r, void* p, size_t n) const { io(fdOf(r), p, n, false, r); }
  template <class T> void sendVec(int r, const std::vector<T>& v) const {
    const uint64_t n = v.size();
    send(r, &n, sizeof n);
    if (n) send(r, v.data(), n * sizeof(T));
  }
  template <class T> std::vector<T> recvVec(int r) const {
    uint64_t n = 0;
    recv(r, &n, sizeof n);
    std::vector<T> v(n);
    if (n) recv(r, v.data(), n * sizeof(T));
    return v;
  }
  void bcast(void* p, size_t n) const {  // rank 0 -> everyone
    if (!multi()) return;
    if (root()) {
      for (int r = 1; r < size; ++r) send(r, p, n);
    } else {
      recv(0, p, n);
    }
  }
  void barrier() const {
    if (!multi()) return;
    char c = 0;
    if (root()) {
      for (int r = 1; r < size; ++r) recv(r, &c, 1);
      for (int r = 1; r < size; ++r) send(r, &c, 1);
    } else {
      send(0, &c, 1);
      recv(0, &c, 1);
    }
  }
  // rank 0: waits for the workers; true when every one of them exited with EXIT_SUCCESS
  bool join() {
    bool ok = true;
    for (pid_t pid : mChildren) {
      int status = 0;
      if (waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) ok = false;
    }
    mChildren.clear();
    return ok;
  }

 private:
  int fdOf(int r) const { return root() ? mFd[r] : mFd[0]; }
  static void io(int fd, void* p, size_t n, bool write, int peer) {
    char* c = static_cast<char*>(p);
    while (n) {
      const ssize_t k = write ? ::send(fd, c, n, MSG_NOSIGNAL) : ::read(fd, c, n);
      if (k < 0 && errno == EINTR) continue;
      if (k <= 0) throw std::runtime_error("Error: the process of rank " + std::to_string(peer) + " of the slab-decomposed run terminated.");
      c += k, n -= (size_t)k;
    }
  }
  std::vector<int> mFd;  // rank 0: [r] = socket to rank r; worker: [0] = socket to rank 0
  std::vector<pid_t> mChildren;
};

}  // namespace kwhost
