// minih5.cpp -- implementation of the HDF5 API subset declared in hdf5.h / hdf5_hl.h over an in-memory object tree that is
// read from / written to REAL HDF5 files (the subset of the file format described above save_hdf5 below): a file is parsed
// completely on H5Fopen and written completely on H5Fclose of a writable file.
//
// Legacy: files in the private "KWH5" container of round 1 are still read (magic "KWH5\x00\x01\x00\x00").
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "hdf5_hl.h"

namespace {

struct Attr {
  int type = 0;  // 0 string, 1 int64, 2 float32
  std::string s;
  long long i = 0;
  float f = 0.f;
};

// a read-only mapping of an opened file; contiguous datasets of a matching element type are served straight from it
struct Mapping {
  void* p = nullptr;
  size_t n = 0;
  ~Mapping() {
    if (p) munmap(p, n);
  }
};

struct Node {
  bool is_group = true;
  int dtype = 1;  // 1 float32, 2 uint64
  std::vector<hsize_t> dims, chunk;
  unsigned deflate = 0;
  bool has_deflate = false;  // the deflate filter is part of the pipeline (the reference registers it even at level 0, Hdf5File.cpp:345)
  std::vector<uint8_t> data;
  // zero-copy view of a contiguous dataset inside the file mapping (until the dataset is written to): a 1024^3 input is ~56 GB,
  // and every rank of a slab-decomposed run opens it -- the pages are shared through the page cache instead of copied per process
  const uint8_t* view = nullptr;
  std::shared_ptr<Mapping> mapping;
  const uint8_t* bytes() const { return view ? view : data.data(); }
  size_t nbytes() const { return view ? elems() * esize() : data.size(); }
  void materialize() {
    if (!view) return;
    data.assign(view, view + elems() * esize());
    view = nullptr;
    mapping.reset();
  }
  std::map<std::string, Attr> attrs;
  std::map<std::string, std::unique_ptr<Node>> children;
  std::vector<std::string> order;  // creation order of children
  size_t elems() const {
    size_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
  size_t esize() const { return dtype == 1 ? 4 : 8; }
};

struct File {
  Node root;
  std::string path;
  bool writable = false;
};

struct Space {
  std::vector<hsize_t> dims, start, count;
  bool selected = false;
};

struct PList {
  std::vector<hsize_t> chunk;
  unsigned deflate = 0;
  bool has_deflate = false;
};

enum Kind { kFile, kGroup, kDataset, kSpace, kPList };
struct Handle {
  Kind kind;
  File* file = nullptr;    // file / group / dataset
  Node* node = nullptr;    // group / dataset (file: root)
  std::string name;        // absolute path
  Space space;
  PList plist;
};

std::mutex g_mu;
std::map<hid_t, std::unique_ptr<Handle>> g_handles;
hid_t g_next = 0x100;

hid_t new_handle(std::unique_ptr<Handle> h) {
  std::lock_guard<std::mutex> lk(g_mu);
  const hid_t id = g_next++;
  g_handles[id] = std::move(h);
  return id;
}
Handle* get(hid_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(id);
  return it == g_handles.end() ? nullptr : it->second.get();
}
void drop(hid_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_handles.erase(id);
}

std::string join(const std::string& base, const std::string& name) {
  if (!name.empty() && name[0] == '/') return name;
  if (base == "/") return "/" + name;
  return base + "/" + name;
}

// resolve `name` (relative or absolute, possibly with several components, "/" or "." allowed) from a location handle
Node* resolve(Handle* loc, const char* name, std::string* abs = nullptr) {
  if (!loc || !loc->node) return nullptr;
  std::string n = name ? name : "";
  Node* cur = loc->node;
  std::string path = loc->name;
  if (!n.empty() && n[0] == '/') cur = &loc->file->root, path = "/";
  size_t pos = 0;
  while (pos < n.size()) {
    size_t next = n.find('/', pos);
    if (next == std::string::npos) next = n.size();
    const std::string part = n.substr(pos, next - pos);
    pos = next + 1;
    if (part.empty() || part == ".") continue;
    auto it = cur->children.find(part);
    if (it == cur->children.end()) return nullptr;
    cur = it->second.get();
    path = join(path, part);
  }
  if (abs) *abs = path;
  return cur;
}

Node* create_child(Handle* loc, const char* name, std::string* abs) {
  // parent = everything before the last component
  std::string n = name;
  std::string parent = ".", leaf = n;
  const size_t slash = n.rfind('/');
  if (slash != std::string::npos) parent = slash == 0 ? "/" : n.substr(0, slash), leaf = n.substr(slash + 1);
  std::string pabs;
  Node* p = resolve(loc, parent.c_str(), &pabs);
  if (!p || !p->is_group || leaf.empty() || p->children.count(leaf)) return nullptr;
  p->children[leaf].reset(new Node());
  p->order.push_back(leaf);
  *abs = join(pabs, leaf);
  return p->children[leaf].get();
}

// ---- serialisation ----------------------------------------------------------------------------------------------------
const char kMagic[8] = {'K', 'W', 'H', '5', 0, 1, 0, 0};

template <class T> void put(FILE* f, T v) { fwrite(&v, sizeof(T), 1, f); }
void put_str(FILE* f, const std::string& s) {
  put<uint32_t>(f, (uint32_t)s.size());
  fwrite(s.data(), 1, s.size(), f);
}
template <class T> bool take(FILE* f, T* v) { return fread(v, sizeof(T), 1, f) == 1; }
bool take_str(FILE* f, std::string* s) {
  uint32_t n;
  if (!take(f, &n)) return false;
  s->resize(n);
  return n == 0 || fread(&(*s)[0], 1, n, f) == n;
}

bool load_kwh5(File* file) {
  FILE* f = fopen(file->path.c_str(), "rb");
  if (!f) return false;
  char magic[8];
  uint64_t nrec = 0;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0 && take(f, &nrec);
  for (uint64_t r = 0; ok && r < nrec; ++r) {
    std::string path;
    uint8_t kind;
    uint32_t nattrs;
    ok = take_str(f, &path) && take(f, &kind) && take(f, &nattrs);
    if (!ok) break;
    Node* n = &file->root;
    if (path != "/") {
      // walk / create
      size_t pos = 1;
      while (pos <= path.size()) {
        size_t next = path.find('/', pos);
        if (next == std::string::npos) next = path.size();
        const std::string part = path.substr(pos, next - pos);
        pos = next + 1;
        auto it = n->children.find(part);
        if (it == n->children.end()) {
          n->children[part].reset(new Node());
          n->order.push_back(part);
          it = n->children.find(part);
        }
        n = it->second.get();
      }
    }
    n->is_group = kind == 0;
    n->dtype = kind == 0 ? 1 : kind;
    for (uint32_t a = 0; ok && a < nattrs; ++a) {
      std::string name;
      uint8_t type;
      ok = take_str(f, &name) && take(f, &type);
      Attr at;
      at.type = type;
      if (ok && type == 0) ok = take_str(f, &at.s);
      else if (ok && type == 1) {
        int64_t v;
        ok = take(f, &v);
        at.i = v;
      } else if (ok) ok = take(f, &at.f);
      if (ok) n->attrs[name] = at;
    }
    if (ok && kind != 0) {
      uint32_t rank, crank;
      ok = take(f, &rank);
      n->dims.resize(ok ? rank : 0);
      for (auto& d : n->dims) {
        uint64_t v;
        ok = ok && take(f, &v);
        d = v;
      }
      ok = ok && take(f, &crank);
      n->chunk.resize(ok ? crank : 0);
      for (auto& d : n->chunk) {
        uint64_t v;
        ok = ok && take(f, &v);
        d = v;
      }
      ok = ok && take(f, &n->deflate);
      if (ok) {
        n->data.resize(n->elems() * n->esize());
        ok = n->data.empty() || fread(n->data.data(), 1, n->data.size(), f) == n->data.size();
      }
    }
  }
  fclose(f);
  return ok;
}

// ---- the HDF5 file format (subset) ---------------------------------------------------------------------------------
// Written against the HDF5 File Format Specification 1.8: superblock version 0, version-1 object headers (continuation
// blocks followed on read), old-style groups (symbol-table message -> version-1 B-tree of group nodes + local heap +
// symbol-table nodes), contiguous / compact / chunked datasets (version-1 chunk B-tree of any depth), the deflate filter
// (zlib), IEEE float32 and 64-bit integers, attributes holding fixed-length strings or float32 / int64 scalars -- exactly
// what Hdf5/Hdf5File.cpp:97-1086 of the reference produces and consumes (libver "earliest", the default of libhdf5 1.8 and
// of MATLAB's h5create / h5write).  tools/h5lite.py is the independent Python implementation the tests cross-check
// against; no libhdf5 exists in this image, so conformance rests on the specification's structure tables.
constexpr uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;
constexpr int kLeafK = 32, kInternalK = 16, kChunkK = 32;
const unsigned char kSig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};

struct Buf {  // little-endian byte builder
  std::vector<uint8_t> b;
  void u8(unsigned v) { b.push_back((uint8_t)v); }
  void u16(unsigned v) { u8(v & 0xff), u8((v >> 8) & 0xff); }
  void u32(uint64_t v) { for (int i = 0; i < 4; ++i) u8((unsigned)(v >> (8 * i)) & 0xff); }
  void u64(uint64_t v) { for (int i = 0; i < 8; ++i) u8((unsigned)(v >> (8 * i)) & 0xff); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void bytes(const void* p, size_t n) { b.insert(b.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
  void append(const Buf& o) { b.insert(b.end(), o.b.begin(), o.b.end()); }
  void pad8() { zeros((8 - b.size() % 8) % 8); }
  size_t size() const { return b.size(); }
};

struct Writer {
  FILE* f;
  uint64_t pos = 0;
  bool ok = true;
  uint64_t alloc(uint64_t n) {
    pos += (8 - pos % 8) % 8;
    const uint64_t a = pos;
    pos += n;
    return a;
  }
  void put(uint64_t addr, const void* p, size_t n) {
    if (!n) return;
    if (fseeko(f, (off_t)addr, SEEK_SET) != 0 || fwrite(p, 1, n, f) != n) ok = false;
  }
  void put(uint64_t addr, const Buf& b) { put(addr, b.b.data(), b.b.size()); }
};

Buf dt_f32() {
  Buf b;
  b.u8(0x11), b.u8(0x20), b.u8(0x1F), b.u8(0), b.u32(4);
  b.u16(0), b.u16(32), b.u8(23), b.u8(8), b.u8(0), b.u8(23), b.u32(127);
  return b;
}
Buf dt_int(int size, bool is_signed) {
  Buf b;
  b.u8(0x10), b.u8(is_signed ? 0x08 : 0x00), b.u8(0), b.u8(0), b.u32(size);
  b.u16(0), b.u16(8 * size);
  return b;
}
Buf dt_str(size_t size) {
  Buf b;
  b.u8(0x13), b.u8(0), b.u8(0), b.u8(0), b.u32(size);
  return b;
}
Buf space_msg(const std::vector<hsize_t>& dims) {
  Buf b;
  b.u8(1), b.u8((unsigned)dims.size()), b.u8(0), b.zeros(5);
  for (auto d : dims) b.u64(d);
  return b;
}
void add_msg(Buf* out, unsigned type, Buf body, unsigned flags = 0) {
  body.pad8();
  out->u16(type), out->u16((unsigned)body.size()), out->u8(flags), out->zeros(3);
  out->append(body);
}
void add_attr(Buf* out, const std::string& name, const Attr& a) {
  Buf raw, dt, sp;
  if (a.type == 0) {
    raw.bytes(a.s.c_str(), a.s.size() + 1);
    dt = dt_str(a.s.size() + 1), sp = space_msg({});
  } else if (a.type == 1) {
    raw.u64((uint64_t)a.i);
    dt = dt_int(8, true), sp = space_msg({1});
  } else {
    uint32_t bits;
    memcpy(&bits, &a.f, 4);
    raw.u32(bits);
    dt = dt_f32(), sp = space_msg({1});
  }
  Buf body;
  body.u8(1), body.u8(0), body.u16((unsigned)name.size() + 1), body.u16((unsigned)dt.size()), body.u16((unsigned)sp.size());
  Buf nm;
  nm.bytes(name.c_str(), name.size() + 1);
  nm.pad8(), dt.pad8(), sp.pad8();
  body.append(nm), body.append(dt), body.append(sp), body.append(raw);
  add_msg(out, 0x0C, body);
}
Buf object_header(const Buf& msgs, int nmsgs) {
  Buf h;
  h.u8(1), h.u8(0), h.u16(nmsgs), h.u32(1), h.u32(msgs.size()), h.zeros(4);
  h.append(msgs);
  return h;
}

struct BtEntry {
  Buf key, end_key;  // the key in front of the child and the key that closes a node ending with this child
  uint64_t child;
};
uint64_t write_btree_node(Writer& w, int type, int level, const std::vector<BtEntry>& e, size_t first, size_t count, size_t key_size, int capacity) {
  Buf n;
  n.bytes("TREE", 4), n.u8(type), n.u8(level), n.u16((unsigned)count), n.u64(kUndef), n.u64(kUndef);
  for (size_t i = first; i < first + count; ++i) n.append(e[i].key), n.u64(e[i].child);
  n.append(e[first + count - 1].end_key);
  const size_t full = 24 + (size_t)capacity * (key_size + 8) + key_size;
  n.zeros(full - n.size());
  const uint64_t addr = w.alloc(full);
  w.put(addr, n);
  return addr;
}

Buf chunk_key(uint32_t size, const std::vector<hsize_t>& off) {
  Buf k;
  k.u32(size), k.u32(0);
  for (auto o : off) k.u64(o);
  k.u64(0);
  return k;
}
// chunks in row-major (= lexicographic) order, then the B-tree bottom-up; returns the root node
uint64_t write_chunks(Writer& w, const Node& ds, bool filtered) {
  const size_t rank = ds.dims.size(), es = ds.esize();
  std::vector<hsize_t> counts(rank), idx(rank, 0), off(rank);
  size_t nchunks = 1, celems = 1;
  for (size_t d = 0; d < rank; ++d) counts[d] = (ds.dims[d] + ds.chunk[d] - 1) / ds.chunk[d], nchunks *= counts[d], celems *= ds.chunk[d];
  std::vector<size_t> dstride(rank), cstride(rank);
  for (size_t d = rank, s = 1, c = 1; d-- > 0;) dstride[d] = s, cstride[d] = c, s *= ds.dims[d], c *= ds.chunk[d];
  std::vector<uint8_t> block(celems * es), packed;
  std::vector<BtEntry> entries(nchunks);
  std::vector<std::vector<hsize_t>> offs(nchunks);
  for (size_t c = 0; c < nchunks; ++c) {
    for (size_t d = 0; d < rank; ++d) off[d] = idx[d] * ds.chunk[d];
    offs[c] = off;
    // gather the (possibly clipped) block; rows along the last dimension are contiguous in both layouts
    std::fill(block.begin(), block.end(), 0);
    std::vector<hsize_t> ext(rank);
    size_t rows = 1;
    for (size_t d = 0; d < rank; ++d) ext[d] = std::min<hsize_t>(ds.chunk[d], ds.dims[d] - off[d]), rows *= d + 1 < rank ? ext[d] : 1;
    std::vector<hsize_t> r(rank, 0);
    for (size_t row = 0; row < rows; ++row) {
      size_t src = 0, dst = 0;
      for (size_t d = 0; d + 1 < rank; ++d) src += (off[d] + r[d]) * dstride[d], dst += r[d] * cstride[d];
      src += off[rank - 1];
      memcpy(block.data() + dst * es, ds.bytes() + src * es, ext[rank - 1] * es);
      for (size_t d = rank - 1; d-- > 0;) {
        if (++r[d] < ext[d]) break;
        r[d] = 0;
      }
    }
    const uint8_t* out = block.data();
    size_t out_size = block.size();
    if (filtered) {
      uLongf cap = compressBound(block.size());
      packed.resize(cap);
      if (compress2(packed.data(), &cap, block.data(), block.size(), (int)ds.deflate) != Z_OK) w.ok = false;
      out = packed.data(), out_size = cap;
    }
    const uint64_t addr = w.alloc(out_size);
    w.put(addr, out, out_size);
    entries[c].key = chunk_key((uint32_t)out_size, off);
    entries[c].child = addr;
    for (size_t d = rank; d-- > 0;) {
      if (++idx[d] < counts[d]) break;
      idx[d] = 0;
    }
  }
  std::vector<hsize_t> end(rank, 0);
  end[0] = counts[0] * ds.chunk[0];
  for (size_t c = 0; c < nchunks; ++c) entries[c].end_key = chunk_key(0, c + 1 < nchunks ? offs[c + 1] : end);
  const size_t key_size = 8 + 8 * (rank + 1);
  for (int level = 0;; ++level) {
    std::vector<BtEntry> up;
    for (size_t i = 0; i < entries.size(); i += 2 * kChunkK) {
      const size_t cnt = std::min<size_t>(2 * kChunkK, entries.size() - i);
      BtEntry e;
      e.child = write_btree_node(w, 1, level, entries, i, cnt, key_size, 2 * kChunkK);
      e.key = entries[i].key, e.end_key = entries[i + cnt - 1].end_key;
      up.push_back(std::move(e));
    }
    if (up.size() == 1) return up[0].child;
    entries.swap(up);
  }
}

uint64_t write_dataset(Writer& w, const Node& ds) {
  Buf msgs;
  int n = 0;
  std::vector<hsize_t> dims = ds.dims;
  add_msg(&msgs, 0x01, space_msg(dims)), ++n;
  add_msg(&msgs, 0x03, ds.dtype == 1 ? dt_f32() : dt_int(8, false), 1), ++n;
  const bool chunked = !ds.chunk.empty() && ds.chunk.size() == ds.dims.size() && ds.elems() > 0;
  Buf fill;
  fill.u8(2), fill.u8(chunked ? 3 : 2), fill.u8(2), fill.u8(0);
  add_msg(&msgs, 0x05, fill), ++n;
  if (chunked) {
    const bool filtered = ds.has_deflate;
    if (filtered) {
      Buf fp;
      fp.u8(1), fp.u8(1), fp.zeros(6);
      fp.u16(1), fp.u16(0), fp.u16(1), fp.u16(1), fp.u32(ds.deflate), fp.zeros(4);
      add_msg(&msgs, 0x0B, fp), ++n;
    }
    const uint64_t root = write_chunks(w, ds, filtered);
    Buf lay;
    lay.u8(3), lay.u8(2), lay.u8((unsigned)ds.dims.size() + 1), lay.u64(root);
    for (auto c : ds.chunk) lay.u32(c);
    lay.u32(ds.esize());
    add_msg(&msgs, 0x08, lay), ++n;
  } else {
    const uint64_t addr = ds.nbytes() == 0 ? kUndef : w.alloc(ds.nbytes());
    w.put(addr, ds.bytes(), ds.nbytes());
    Buf lay;
    lay.u8(3), lay.u8(1), lay.u64(addr), lay.u64(ds.nbytes());
    add_msg(&msgs, 0x08, lay), ++n;
  }
  for (auto& kv : ds.attrs) add_attr(&msgs, kv.first, kv.second), ++n;
  const Buf h = object_header(msgs, n);
  const uint64_t addr = w.alloc(h.size());
  w.put(addr, h);
  return addr;
}

struct GroupAddr {
  uint64_t header, btree, heap;
};
GroupAddr write_group(Writer& w, const Node& g) {
  struct Ent {
    std::string name;
    uint64_t header;
    uint32_t cache;
    uint64_t btree, heap;
    uint64_t name_off;
  };
  std::vector<Ent> ents;
  for (auto& kv : g.children) {  // std::map iterates in strcmp order, the order symbol-table entries must have
    Ent e{kv.first, 0, 0, 0, 0, 0};
    if (kv.second->is_group) {
      const GroupAddr a = write_group(w, *kv.second);
      e.header = a.header, e.cache = 1, e.btree = a.btree, e.heap = a.heap;
    } else {
      e.header = write_dataset(w, *kv.second);
    }
    ents.push_back(e);
  }
  Buf heap;
  heap.zeros(8);  // offset 0: the empty string, first key of the B-tree
  for (auto& e : ents) {
    e.name_off = heap.size();
    heap.bytes(e.name.c_str(), e.name.size() + 1);
    heap.pad8();
  }
  const uint64_t free_off = heap.size();
  heap.u64(1), heap.u64(16);  // one 16-byte free block closes the segment; next = 1 ends the free list
  const uint64_t heap_data = w.alloc(heap.size());
  w.put(heap_data, heap);
  Buf hh;
  hh.bytes("HEAP", 4), hh.u8(0), hh.zeros(3), hh.u64(heap.size()), hh.u64(free_off), hh.u64(heap_data);
  const uint64_t heap_addr = w.alloc(hh.size());
  w.put(heap_addr, hh);
  std::vector<BtEntry> nodes;
  for (size_t i = 0; i == 0 || i < ents.size(); i += 2 * kLeafK) {
    const size_t cnt = std::min<size_t>(2 * kLeafK, ents.size() - std::min(i, ents.size()));
    Buf sn;
    sn.bytes("SNOD", 4), sn.u8(1), sn.u8(0), sn.u16((unsigned)cnt);
    for (size_t j = i; j < i + cnt; ++j) {
      sn.u64(ents[j].name_off), sn.u64(ents[j].header), sn.u32(ents[j].cache), sn.u32(0);
      if (ents[j].cache) sn.u64(ents[j].btree), sn.u64(ents[j].heap);
      else sn.zeros(16);
    }
    sn.zeros(8 + 2 * kLeafK * 40 - sn.size());
    BtEntry e;
    e.child = w.alloc(sn.size());
    w.put(e.child, sn);
    e.key.u64(nodes.empty() ? 0 : ents[i - 1].name_off);  // key i = the largest name of child i - 1
    e.end_key.u64(cnt ? ents[i + cnt - 1].name_off : 0);
    nodes.push_back(std::move(e));
  }
  if (nodes.size() > (size_t)2 * kInternalK) w.ok = false;  // > 2048 members in one group: not produced by this code base
  GroupAddr a;
  a.heap = heap_addr;
  a.btree = write_btree_node(w, 0, 0, nodes, 0, std::min<size_t>(nodes.size(), 2 * kInternalK), 8, 2 * kInternalK);
  Buf msgs;
  int n = 0;
  Buf st;
  st.u64(a.btree), st.u64(a.heap);
  add_msg(&msgs, 0x11, st), ++n;
  for (auto& kv : g.attrs) add_attr(&msgs, kv.first, kv.second), ++n;
  const Buf h = object_header(msgs, n);
  a.header = w.alloc(h.size());
  w.put(a.header, h);
  return a;
}

bool save_hdf5(const File& file) {
  // written beside the target and renamed over it: datasets of this very file may still be served from its mapping
  const std::string tmp = file.path + ".minih5-tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  Writer w{f};
  w.alloc(96);
  const GroupAddr root = write_group(w, file.root);
  const uint64_t eof = w.alloc(0);
  Buf sb;
  sb.bytes(kSig, 8);
  sb.u8(0), sb.u8(0), sb.u8(0), sb.u8(0), sb.u8(0), sb.u8(8), sb.u8(8), sb.u8(0);
  sb.u16(kLeafK), sb.u16(kInternalK), sb.u32(0);
  sb.u64(0), sb.u64(kUndef), sb.u64(eof), sb.u64(kUndef);
  sb.u64(0), sb.u64(root.header), sb.u32(1), sb.u32(0), sb.u64(root.btree), sb.u64(root.heap);
  w.put(0, sb);
  bool ok = w.ok && fflush(f) == 0;
  if (ok) {  // the file ends at the end-of-file address even when the last allocation was padding
    fseeko(f, 0, SEEK_END);
    const uint64_t have = (uint64_t)ftello(f);
    if (have < eof) {
      const char z = 0;
      ok = fseeko(f, (off_t)eof - 1, SEEK_SET) == 0 && fwrite(&z, 1, 1, f) == 1;
    }
  }
  ok = fclose(f) == 0 && ok;
  if (ok) ok = rename(tmp.c_str(), file.path.c_str()) == 0;
  else remove(tmp.c_str());
  return ok;
}

// ---- reader ----------------------------------------------------------------------------------------------------------------
struct Reader {
  const uint8_t* b = nullptr;
  size_t n = 0;
  std::shared_ptr<Mapping> mapping;
  std::string error;
  bool fail(const std::string& m) {
    if (error.empty()) error = m;
    return false;
  }
  bool in(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
  uint64_t u(uint64_t off, int bytes) const {
    uint64_t v = 0;
    if (!in(off, bytes)) return 0;
    for (int i = 0; i < bytes; ++i) v |= (uint64_t)b[off + i] << (8 * i);
    return v;
  }
};
struct Msg {
  unsigned type, flags;
  uint64_t off, size;  // body
};
bool read_messages(Reader& r, uint64_t addr, std::vector<Msg>* out) {
  if (!r.in(addr, 16)) return r.fail("object header outside the file");
  if (r.b[addr] != 1) return r.fail("object header version " + std::to_string(r.b[addr]) + " is not supported (only version 1: files written with libver 'earliest')");
  const unsigned nmsg = (unsigned)r.u(addr + 2, 2);
  std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, r.u(addr + 8, 4)}};
  for (size_t bi = 0; bi < blocks.size() && out->size() < nmsg; ++bi) {
    uint64_t p = blocks[bi].first, left = blocks[bi].second;
    if (!r.in(p, left)) return r.fail("object header block outside the file");
    while (left >= 8 && out->size() < nmsg) {
      Msg m{(unsigned)r.u(p, 2), r.b[p + 4], p + 8, r.u(p + 2, 2)};
      if (m.size > left - 8) return r.fail("object header message overruns its block");
      if (m.type == 0x10) blocks.push_back({r.u(m.off, 8), r.u(m.off + 8, 8)});
      out->push_back(m);
      p += 8 + m.size, left -= 8 + m.size;
    }
  }
  return true;
}
struct DType {
  char cls = 0;  // 'f', 'i', 'u', 's'
  uint32_t size = 0;
};
bool parse_dtype(Reader& r, uint64_t off, DType* t) {
  const unsigned cls = r.b[off] & 0x0F;
  t->size = (uint32_t)r.u(off + 4, 4);
  if (cls == 1) t->cls = 'f';
  else if (cls == 0) t->cls = (r.b[off + 1] & 0x08) ? 'i' : 'u';
  else if (cls == 3) t->cls = 's';
  else return r.fail("datatype class " + std::to_string(cls) + " is not supported");
  if ((cls == 0 || cls == 1) && (r.b[off + 1] & 1)) return r.fail("big-endian data is not supported");
  return true;
}
bool parse_space(Reader& r, uint64_t off, std::vector<hsize_t>* dims) {
  const unsigned ver = r.b[off], rank = r.b[off + 1];
  if (ver != 1 && ver != 2) return r.fail("dataspace message version " + std::to_string(ver));
  const uint64_t p = off + (ver == 1 ? 8 : 4);
  dims->resize(rank);
  for (unsigned i = 0; i < rank; ++i) (*dims)[i] = r.u(p + 8 * i, 8);
  return true;
}
bool read_attr(Reader& r, const Msg& m, Node* node) {
  const uint64_t o = m.off;
  const unsigned ver = r.b[o];
  const uint64_t ns = r.u(o + 2, 2), ds = r.u(o + 4, 2), ss = r.u(o + 6, 2);
  if (ver < 1 || ver > 3) return r.fail("attribute message version " + std::to_string(ver));
  auto pad = [&](uint64_t n) { return ver == 1 ? (n + 7) / 8 * 8 : n; };
  uint64_t p = o + (ver == 3 ? 9 : 8);
  const std::string name((const char*)r.b + p, strnlen((const char*)r.b + p, ns));
  p += pad(ns);
  DType t;
  if (!parse_dtype(r, p, &t)) return false;
  p += pad(ds);
  std::vector<hsize_t> dims;
  if (!parse_space(r, p, &dims)) return false;
  p += pad(ss);
  Attr a;
  if (t.cls == 's') {
    a.type = 0;
    a.s.assign((const char*)r.b + p, strnlen((const char*)r.b + p, t.size));
  } else if (t.cls == 'f') {
    a.type = 2;
    if (t.size == 4) {
      const uint32_t bits = (uint32_t)r.u(p, 4);
      memcpy(&a.f, &bits, 4);
    } else {
      const uint64_t bits = r.u(p, 8);
      double d;
      memcpy(&d, &bits, 8);
      a.f = (float)d;
    }
  } else {
    a.type = 1;
    uint64_t v = r.u(p, (int)t.size);
    if (t.cls == 'i' && t.size < 8 && (v >> (8 * t.size - 1))) v |= ~0ull << (8 * t.size);
    a.i = (long long)v;
  }
  node->attrs[name] = a;
  return true;
}
struct ChunkLeaf {
  std::vector<hsize_t> off;
  uint32_t size, mask;
  uint64_t addr;
};
bool chunk_leaves(Reader& r, uint64_t addr, size_t rank, std::vector<ChunkLeaf>* out, int depth = 0) {
  if (depth > 16 || !r.in(addr, 24) || memcmp(r.b + addr, "TREE", 4) != 0 || r.b[addr + 4] != 1) return r.fail("chunk B-tree node expected");
  const unsigned level = r.b[addr + 5], n = (unsigned)r.u(addr + 6, 2);
  const uint64_t ks = 8 + 8 * (rank + 1);
  uint64_t p = addr + 24;
  if (!r.in(p, (uint64_t)n * (ks + 8))) return r.fail("chunk B-tree node overruns the file");
  for (unsigned i = 0; i < n; ++i, p += ks + 8) {
    const uint64_t child = r.u(p + ks, 8);
    if (level) {
      if (!chunk_leaves(r, child, rank, out, depth + 1)) return false;
      continue;
    }
    ChunkLeaf l;
    l.size = (uint32_t)r.u(p, 4), l.mask = (uint32_t)r.u(p + 4, 4), l.addr = child;
    l.off.resize(rank);
    for (size_t d = 0; d < rank; ++d) l.off[d] = r.u(p + 8 + 8 * d, 8);
    out->push_back(std::move(l));
  }
  return true;
}
// converts `count` stored elements of type t into the node's element type (float32 or uint64)
bool convert_elems(Reader& r, const DType& t, const uint8_t* src, size_t count, Node* ds, uint8_t* dst) {
  if ((t.cls == 'f' && t.size == 4 && ds->dtype == 1) || (t.cls != 'f' && t.size == 8 && ds->dtype == 2)) {
    memcpy(dst, src, count * t.size);
    return true;
  }
  for (size_t i = 0; i < count; ++i) {
    if (t.cls == 'f' && t.size == 8) {
      double d;
      memcpy(&d, src + 8 * i, 8);
      const float f = (float)d;
      memcpy(dst + 4 * i, &f, 4);
    } else if (t.cls != 'f' && (t.size == 4 || t.size == 2 || t.size == 1)) {
      uint64_t v = 0;
      memcpy(&v, src + t.size * i, t.size);
      if (t.cls == 'i' && (v >> (8 * t.size - 1))) v |= ~0ull << (8 * t.size);
      memcpy(dst + 8 * i, &v, 8);
    } else {
      return r.fail("unsupported element type (float32 / float64 / integers only)");
    }
  }
  return true;
}
bool read_dataset(Reader& r, const std::vector<Msg>& msgs, Node* ds) {
  DType t;
  bool have_space = false, have_type = false;
  const Msg* layout = nullptr;
  bool deflate = false;
  ds->is_group = false;
  for (auto& m : msgs) {
    if (m.type == 0x01) have_space = parse_space(r, m.off, &ds->dims);
    else if (m.type == 0x03) have_type = parse_dtype(r, m.off, &t);
    else if (m.type == 0x08) layout = &m;
    else if (m.type == 0x0C) {
      if (!read_attr(r, m, ds)) return false;
    } else if (m.type == 0x0B) {
      const unsigned ver = r.b[m.off], nf = r.b[m.off + 1];
      uint64_t p = m.off + (ver == 1 ? 8 : 2);
      for (unsigned i = 0; i < nf; ++i) {
        const unsigned id = (unsigned)r.u(p, 2);
        unsigned ncd;
        if (ver == 1 || id >= 256) {
          const uint64_t nl = r.u(p + 2, 2);
          ncd = (unsigned)r.u(p + 6, 2);
          p += 8 + (ver == 1 ? (nl + 7) / 8 * 8 : nl);
        } else {
          ncd = (unsigned)r.u(p + 4, 2);
          p += 6;
        }
        if (id != 1) return r.fail("filter " + std::to_string(id) + " is not supported (only deflate)");
        deflate = true;
        ds->has_deflate = true;
        ds->deflate = ncd ? (unsigned)r.u(p, 4) : 0;
        p += 4 * ncd + ((ver == 1 && ncd % 2) ? 4 : 0);
      }
    }
  }
  if (!r.error.empty()) return false;
  if (!have_space || !have_type || !layout) return r.fail("dataset without dataspace / datatype / layout message");
  if (t.cls == 's') return r.fail("string datasets are not supported");
  ds->dtype = t.cls == 'f' ? 1 : 2;
  const size_t n = ds->elems();
  const uint64_t lo = layout->off;
  if (r.b[lo] != 3) return r.fail("data layout message version " + std::to_string(r.b[lo]) + " is not supported");
  const unsigned cls = r.b[lo + 1];
  const bool same_type = (t.cls == 'f' && t.size == 4) || (t.cls != 'f' && t.size == 8);
  if (cls == 1 && same_type && r.u(lo + 2, 8) != kUndef && n * t.size >= (1u << 16)) {  // large contiguous data: served from the mapping
    const uint64_t addr = r.u(lo + 2, 8);
    if (!r.in(addr, n * t.size)) return r.fail("contiguous data outside the file");
    ds->view = r.b + addr, ds->mapping = r.mapping;
    return true;
  }
  ds->data.assign(n * ds->esize(), 0);
  if (cls == 1) {
    const uint64_t addr = r.u(lo + 2, 8);
    if (addr != kUndef) {
      if (!r.in(addr, n * t.size)) return r.fail("contiguous data outside the file");
      if (!convert_elems(r, t, r.b + addr, n, ds, ds->data.data())) return false;
    }
  } else if (cls == 0) {
    if (!convert_elems(r, t, r.b + lo + 4, n, ds, ds->data.data())) return false;
  } else if (cls == 2) {
    const unsigned nd = r.b[lo + 2];
    const uint64_t root = r.u(lo + 3, 8);
    const size_t rank = nd - 1;
    if (rank != ds->dims.size()) return r.fail("chunk rank differs from the dataspace rank");
    ds->chunk.resize(rank);
    size_t celems = 1;
    for (size_t d = 0; d < rank; ++d) ds->chunk[d] = r.u(lo + 11 + 4 * d, 4), celems *= ds->chunk[d];
    std::vector<ChunkLeaf> leaves;
    if (root != kUndef && !chunk_leaves(r, root, rank, &leaves)) return false;
    std::vector<size_t> dstride(rank), cstride(rank);
    for (size_t d = rank, s = 1, c = 1; d-- > 0;) dstride[d] = s, cstride[d] = c, s *= ds->dims[d], c *= ds->chunk[d];
    std::vector<uint8_t> raw(celems * t.size), conv(celems * ds->esize());
    for (auto& l : leaves) {
      if (!r.in(l.addr, l.size)) return r.fail("chunk outside the file");
      const uint8_t* src = r.b + l.addr;
      if (deflate && !(l.mask & 1)) {
        uLongf got = raw.size();
        if (uncompress(raw.data(), &got, src, l.size) != Z_OK || got != raw.size()) return r.fail("deflate: corrupt chunk");
        src = raw.data();
      } else if (l.size < raw.size()) {
        return r.fail("short chunk");
      }
      if (!convert_elems(r, t, src, celems, ds, conv.data())) return false;
      const size_t es = ds->esize();
      std::vector<hsize_t> ext(rank), rr(rank, 0);
      size_t rows = 1;
      bool inside = true;
      for (size_t d = 0; d < rank; ++d) {
        if (l.off[d] >= ds->dims[d]) inside = false;
        ext[d] = inside ? std::min<hsize_t>(ds->chunk[d], ds->dims[d] - l.off[d]) : 0;
        rows *= d + 1 < rank ? ext[d] : 1;
      }
      if (!inside) continue;
      for (size_t row = 0; row < rows; ++row) {
        size_t dst = 0, s = 0;
        for (size_t d = 0; d + 1 < rank; ++d) dst += (l.off[d] + rr[d]) * dstride[d], s += rr[d] * cstride[d];
        dst += l.off[rank - 1];
        memcpy(ds->data.data() + dst * es, conv.data() + s * es, ext[rank - 1] * es);
        for (size_t d = rank - 1; d-- > 0;) {
          if (++rr[d] < ext[d]) break;
          rr[d] = 0;
        }
      }
    }
  } else {
    return r.fail("layout class " + std::to_string(cls));
  }
  return true;
}
bool read_object(Reader& r, uint64_t addr, Node* node, int depth);
bool read_group_members(Reader& r, uint64_t btree, uint64_t heap, Node* g, int depth) {
  if (!r.in(heap, 32) || memcmp(r.b + heap, "HEAP", 4) != 0) return r.fail("local heap expected");
  const uint64_t hdata = r.u(heap + 24, 8), hsize = r.u(heap + 8, 8);
  if (!r.in(hdata, hsize)) return r.fail("local heap data outside the file");
  std::vector<std::pair<uint64_t, int>> stack{{btree, 0}};
  while (!stack.empty()) {
    const uint64_t addr = stack.back().first;
    const int d = stack.back().second;
    stack.pop_back();
    if (d > 16 || !r.in(addr, 24) || memcmp(r.b + addr, "TREE", 4) != 0 || r.b[addr + 4] != 0) return r.fail("group B-tree node expected");
    const unsigned level = r.b[addr + 5], n = (unsigned)r.u(addr + 6, 2);
    std::vector<uint64_t> kids(n);
    for (unsigned i = 0; i < n; ++i) kids[i] = r.u(addr + 24 + 8 + 16 * i, 8);
    if (level) {
      for (unsigned i = n; i-- > 0;) stack.push_back({kids[i], d + 1});
      continue;
    }
    for (uint64_t sn : kids) {
      if (!r.in(sn, 8) || memcmp(r.b + sn, "SNOD", 4) != 0) return r.fail("symbol table node expected");
      const unsigned ns = (unsigned)r.u(sn + 6, 2);
      for (unsigned i = 0; i < ns; ++i) {
        const uint64_t e = sn + 8 + 40 * i, noff = r.u(e, 8), haddr = r.u(e + 8, 8);
        if (noff >= hsize) return r.fail("link name outside the local heap");
        const std::string name((const char*)r.b + hdata + noff, strnlen((const char*)r.b + hdata + noff, hsize - noff));
        auto& child = g->children[name];
        child.reset(new Node());
        g->order.push_back(name);
        if (!read_object(r, haddr, child.get(), depth + 1)) return false;
      }
    }
  }
  return true;
}
bool read_object(Reader& r, uint64_t addr, Node* node, int depth) {
  if (depth > 32) return r.fail("group nesting too deep");
  std::vector<Msg> msgs;
  if (!read_messages(r, addr, &msgs)) return false;
  for (auto& m : msgs)
    if (m.type == 0x11) {
      node->is_group = true;
      for (auto& a : msgs)
        if (a.type == 0x0C && !read_attr(r, a, node)) return false;
      return read_group_members(r, r.u(m.off, 8), r.u(m.off + 8, 8), node, depth);
    }
  for (auto& m : msgs)
    if (m.type == 0x02 || m.type == 0x06) return r.fail("new-style groups (link messages) are not supported: write the file with libver 'earliest' (the default)");
  return read_dataset(r, msgs, node);
}

bool load_hdf5(File* file, std::string* why) {
  const int fd = open(file->path.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 96) {
    close(fd);
    return false;
  }
  void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return false;
  Reader r;
  r.mapping = std::make_shared<Mapping>();
  r.mapping->p = map, r.mapping->n = (size_t)st.st_size;
  r.b = (const uint8_t*)map, r.n = (size_t)st.st_size;
  bool ok = memcmp(r.b, kSig, 8) == 0 || r.fail("not an HDF5 file");
  if (ok) {
    const unsigned ver = r.b[8];
    if (ver > 1) ok = r.fail("superblock version " + std::to_string(ver) + " is not supported (versions 0 and 1: libver 'earliest')");
    else if (r.b[13] != 8 || r.b[14] != 8) ok = r.fail("only 8-byte offsets and lengths are supported");
    else {
      const uint64_t p = 24 + (ver == 1 ? 4 : 0);
      if (r.u(p, 8) != 0) ok = r.fail("non-zero base address");
      else ok = read_object(r, r.u(p + 32 + 8, 8), &file->root, 0);
    }
  }
  if (!ok && why) *why = r.error;
  return ok;  // the mapping lives as long as a dataset views it
}

bool save(const File& file) { return save_hdf5(file); }
bool load(File* file) {
  unsigned char magic[8] = {};
  FILE* f = fopen(file->path.c_str(), "rb");
  if (!f) return false;
  const bool got = fread(magic, 1, 8, f) == 8;
  fclose(f);
  if (got && memcmp(magic, kMagic, 8) == 0) return load_kwh5(file);
  std::string why;
  if (load_hdf5(file, &why)) return true;
  if (!why.empty() && getenv("MINIH5_VERBOSE")) fprintf(stderr, "minih5: %s: %s\n", file->path.c_str(), why.c_str());
  return false;
}

// ---- selections ---------------------------------------------------------------------------------------------------------
// enumerate the selected region of a space as runs of contiguous elements: calls f(linear_offset, run_length)
template <class F> void for_runs(const std::vector<hsize_t>& dims, const std::vector<hsize_t>& start, const std::vector<hsize_t>& count, F&& f) {
  const int rank = (int)dims.size();
  if (rank == 0) {
    f((size_t)0, (size_t)1);
    return;
  }
  std::vector<size_t> stride(rank, 1);
  for (int d = rank - 2; d >= 0; --d) stride[d] = stride[d + 1] * dims[d + 1];
  // merge trailing dimensions that are fully selected into one run
  int rd = rank - 1;
  size_t run = count[rd];
  while (rd > 0 && start[rd] == 0 && count[rd] == dims[rd]) {
    --rd;
    run = count[rd] * stride[rd];
  }
  std::vector<hsize_t> idx(rank, 0);
  for (;;) {
    size_t off = 0;
    for (int d = 0; d <= rd; ++d) off += (start[d] + (d < rd ? idx[d] : 0)) * stride[d];
    f(off, run);
    int d = rd - 1;
    for (; d >= 0; --d) {
      if (++idx[d] < count[d]) break;
      idx[d] = 0;
    }
    if (d < 0) break;
  }
}

struct Sel {
  std::vector<hsize_t> dims, start, count;
  size_t total() const {
    size_t n = 1;
    for (auto c : count) n *= c;
    return n;
  }
};
Sel selection_of(const Space* sp, const std::vector<hsize_t>& fallback_dims) {
  Sel s;
  s.dims = sp ? sp->dims : fallback_dims;
  if (sp && sp->selected) s.start = sp->start, s.count = sp->count;
  else s.start.assign(s.dims.size(), 0), s.count = s.dims;
  return s;
}

// copy between a dataset (file selection) and a memory buffer (memory selection); to_file: memory -> dataset
herr_t transfer(Node* ds, hid_t mem_type, hid_t mem_space, hid_t file_space, void* buf, bool to_file) {
  if (!ds || ds->is_group) return -1;
  const int want = mem_type == H5T_NATIVE_FLOAT ? 1 : mem_type == H5T_STD_U64LE ? 2 : 0;
  if (want != ds->dtype) return -1;  // no type conversion in this subset
  const size_t es = ds->esize();
  Handle* fs = file_space == H5S_ALL ? nullptr : get(file_space);
  Handle* ms = mem_space == H5S_ALL ? nullptr : get(mem_space);
  Sel fsel = selection_of(fs ? &fs->space : nullptr, ds->dims);
  Sel msel = ms ? selection_of(&ms->space, ds->dims) : Sel();
  if (!ms) {  // memory is a dense buffer holding exactly the selected elements
    msel.dims = {fsel.total()};
    msel.start = {0};
    msel.count = msel.dims;
  }
  if (fsel.total() != msel.total()) return -1;
  // flatten both selections into run lists, then zip
  std::vector<std::pair<size_t, size_t>> fr, mr;
  for_runs(fsel.dims, fsel.start, fsel.count, [&](size_t o, size_t n) { fr.emplace_back(o, n); });
  for_runs(msel.dims, msel.start, msel.count, [&](size_t o, size_t n) { mr.emplace_back(o, n); });
  size_t fi = 0, mi = 0, fo = 0, mo = 0;
  uint8_t* mem = static_cast<uint8_t*>(buf);
  if (to_file) ds->materialize();  // a dataset served from the file mapping becomes an owned buffer on its first write
  const uint8_t* fbase = ds->bytes();
  const size_t fbytes = ds->nbytes();
  while (fi < fr.size() && mi < mr.size()) {
    const size_t n = std::min(fr[fi].second - fo, mr[mi].second - mo);
    const size_t foff = (fr[fi].first + fo) * es;
    uint8_t* mp = mem + (mr[mi].first + mo) * es;
    if (foff + n * es > fbytes) return -1;
    if (to_file) memcpy(ds->data.data() + foff, mp, n * es);
    else memcpy(mp, fbase + foff, n * es);
    fo += n, mo += n;
    if (fo == fr[fi].second) ++fi, fo = 0;
    if (mo == mr[mi].second) ++mi, mo = 0;
  }
  return 0;
}

Attr* find_attr(hid_t loc, const char* obj, const char* attr, bool create) {
  Node* n = resolve(get(loc), obj);
  if (!n) return nullptr;
  auto it = n->attrs.find(attr);
  if (it == n->attrs.end()) {
    if (!create) return nullptr;
    it = n->attrs.emplace(attr, Attr()).first;
  }
  return &it->second;
}

}  // namespace

extern "C" {

herr_t H5Eset_auto(hid_t, H5E_auto_t, void*) { return 0; }

hid_t H5Fcreate(const char* name, unsigned flags, hid_t, hid_t) {
  if ((flags & H5F_ACC_EXCL) && access(name, F_OK) == 0) return -1;
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kFile;
  h->file = new File();
  h->file->path = name;
  h->file->writable = true;
  h->node = &h->file->root;
  h->name = "/";
  if (!save(*h->file)) {  // fail early when the path is not writable, like H5Fcreate
    delete h->file;
    return -1;
  }
  return new_handle(std::move(h));
}
hid_t H5Fopen(const char* name, unsigned flags, hid_t) {
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kFile;
  h->file = new File();
  h->file->path = name;
  h->file->writable = (flags & H5F_ACC_RDWR) != 0;
  if (!load(h->file)) {
    delete h->file;
    return -1;
  }
  h->node = &h->file->root;
  h->name = "/";
  return new_handle(std::move(h));
}
herr_t H5Fclose(hid_t file) {
  Handle* h = get(file);
  if (!h || h->kind != kFile) return -1;
  herr_t rc = 0;
  if (h->file->writable && !save(*h->file)) rc = -1;
  delete h->file;
  drop(file);
  return rc;
}
htri_t H5Fis_hdf5(const char* name) {
  FILE* f = fopen(name, "rb");
  if (!f) return -1;
  char magic[8];
  const bool ok = fread(magic, 1, 8, f) == 8 && (memcmp(magic, kSig, 8) == 0 || memcmp(magic, kMagic, 8) == 0);
  fclose(f);
  return ok ? 1 : 0;
}
herr_t H5Fget_filesize(hid_t file, hsize_t* size) {
  Handle* h = get(file);
  if (!h || h->kind != kFile) return -1;
  // size the file would have on disk now
  struct Acc {
    static hsize_t of(const Node& n) {
      hsize_t s = 64 + n.nbytes();
      for (auto& kv : n.children) s += of(*kv.second);
      return s;
    }
  };
  *size = Acc::of(h->file->root);
  return 0;
}

hid_t H5Gcreate(hid_t loc, const char* name, hid_t, hid_t, hid_t) {
  Handle* l = get(loc);
  std::string abs;
  Node* n = l ? create_child(l, name, &abs) : nullptr;
  if (!n) return -1;
  n->is_group = true;
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kGroup, h->file = l->file, h->node = n, h->name = abs;
  return new_handle(std::move(h));
}
hid_t H5Gopen(hid_t loc, const char* name, hid_t) {
  Handle* l = get(loc);
  std::string abs;
  Node* n = resolve(l, name, &abs);
  if (!n || !n->is_group) return -1;
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kGroup, h->file = l->file, h->node = n, h->name = abs;
  return new_handle(std::move(h));
}
herr_t H5Gclose(hid_t group) {
  drop(group);
  return 0;
}

htri_t H5Lexists(hid_t loc, const char* name, hid_t) { return resolve(get(loc), name) ? 1 : 0; }
htri_t H5Oexists_by_name(hid_t loc, const char* name, hid_t) { return resolve(get(loc), name) ? 1 : 0; }
herr_t H5Oget_info_by_name(hid_t loc, const char* name, H5O_info_t* info, hid_t) {
  Node* n = resolve(get(loc), name);
  if (!n) return -1;
  info->type = n->is_group ? H5O_TYPE_GROUP : H5O_TYPE_DATASET;
  return 0;
}
ssize_t H5Iget_name(hid_t id, char* name, size_t size) {
  Handle* h = get(id);
  if (!h) return -1;
  if (name && size) {
    strncpy(name, h->name.c_str(), size);
    name[size - 1] = 0;
  }
  return (ssize_t)h->name.size();
}

hid_t H5Dcreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t, hid_t dcpl, hid_t) {
  Handle* l = get(loc);
  Handle* sp = get(space);
  if (!l || !sp || sp->kind != kSpace) return -1;
  if (type != H5T_NATIVE_FLOAT && type != H5T_STD_U64LE) return -1;
  std::string abs;
  Node* n = create_child(l, name, &abs);
  if (!n) return -1;
  n->is_group = false;
  n->dtype = type == H5T_NATIVE_FLOAT ? 1 : 2;
  n->dims = sp->space.dims;
  if (Handle* pl = dcpl ? get(dcpl) : nullptr) n->chunk = pl->plist.chunk, n->deflate = pl->plist.deflate, n->has_deflate = pl->plist.has_deflate;
  n->data.assign(n->elems() * n->esize(), 0);
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kDataset, h->file = l->file, h->node = n, h->name = abs;
  return new_handle(std::move(h));
}
hid_t H5Dopen(hid_t loc, const char* name, hid_t) {
  Handle* l = get(loc);
  std::string abs;
  Node* n = resolve(l, name, &abs);
  if (!n || n->is_group) return -1;
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kDataset, h->file = l->file, h->node = n, h->name = abs;
  return new_handle(std::move(h));
}
herr_t H5Dclose(hid_t dset) {
  drop(dset);
  return 0;
}
hid_t H5Dget_space(hid_t dset) {
  Handle* d = get(dset);
  if (!d || d->kind != kDataset) return -1;
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kSpace;
  h->space.dims = d->node->dims;
  return new_handle(std::move(h));
}
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t, void* buf) {
  Handle* d = get(dset);
  return d ? transfer(d->node, mem_type, mem_space, file_space, buf, false) : -1;
}
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t, const void* buf) {
  Handle* d = get(dset);
  if (!d || !d->file->writable) return -1;
  return transfer(d->node, mem_type, mem_space, file_space, const_cast<void*>(buf), true);
}

hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t*) {
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kSpace;
  h->space.dims.assign(dims, dims + rank);
  return new_handle(std::move(h));
}
herr_t H5Sclose(hid_t space) {
  drop(space);
  return 0;
}
int H5Sget_simple_extent_ndims(hid_t space) {
  Handle* h = get(space);
  return h ? (int)h->space.dims.size() : -1;
}
herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t, const hsize_t* start, const hsize_t* stride, const hsize_t* count, const hsize_t* block) {
  Handle* h = get(space);
  if (!h || h->kind != kSpace || stride || block) return -1;  // unit stride and block only
  const size_t r = h->space.dims.size();
  h->space.start.assign(start, start + r);
  h->space.count.assign(count, count + r);
  for (size_t d = 0; d < r; ++d)
    if (start[d] + count[d] > h->space.dims[d]) return -1;
  h->space.selected = true;
  return 0;
}
herr_t H5Sselect_elements(hid_t, H5S_seloper_t, size_t, const hsize_t*) { return -1; }  // only reached from dead code (Hdf5File.cpp:637-688)

hid_t H5Pcreate(hid_t) {
  std::unique_ptr<Handle> h(new Handle());
  h->kind = kPList;
  return new_handle(std::move(h));
}
herr_t H5Pclose(hid_t plist) {
  drop(plist);
  return 0;
}
herr_t H5Pset_chunk(hid_t plist, int ndims, const hsize_t* dim) {
  Handle* h = get(plist);
  if (!h) return -1;
  h->plist.chunk.assign(dim, dim + ndims);
  return 0;
}
herr_t H5Pset_deflate(hid_t plist, unsigned level) {
  Handle* h = get(plist);
  if (!h) return -1;
  h->plist.deflate = level;
  h->plist.has_deflate = true;
  return 0;
}

// ---- lite API -------------------------------------------------------------------------------------------------------
herr_t H5LTread_dataset(hid_t loc, const char* name, hid_t type, void* buffer) {
  Node* n = resolve(get(loc), name);
  return n ? transfer(n, type, H5S_ALL, H5S_ALL, buffer, false) : -1;
}
herr_t H5LTget_dataset_info(hid_t loc, const char* name, hsize_t* dims, H5T_class_t* cls, size_t* type_size) {
  Node* n = resolve(get(loc), name);
  if (!n || n->is_group) return -1;
  if (dims)
    for (size_t d = 0; d < n->dims.size(); ++d) dims[d] = n->dims[d];
  if (cls) *cls = n->dtype;
  if (type_size) *type_size = n->esize();
  return 0;
}
herr_t H5LTget_dataset_ndims(hid_t loc, const char* name, int* rank) {
  Node* n = resolve(get(loc), name);
  if (!n || n->is_group) return -1;
  *rank = (int)n->dims.size();
  return 0;
}
herr_t H5LTfind_dataset(hid_t loc, const char* name) {
  Node* n = resolve(get(loc), name);
  return n && !n->is_group ? 1 : 0;
}
herr_t H5LTset_attribute_string(hid_t loc, const char* obj, const char* attr, const char* value) {
  Attr* a = find_attr(loc, obj, attr, true);
  if (!a) return -1;
  a->type = 0, a->s = value;
  return 0;
}
herr_t H5LTset_attribute_long_long(hid_t loc, const char* obj, const char* attr, const long long* value, size_t) {
  Attr* a = find_attr(loc, obj, attr, true);
  if (!a) return -1;
  a->type = 1, a->i = *value;
  return 0;
}
herr_t H5LTset_attribute_float(hid_t loc, const char* obj, const char* attr, const float* value, size_t) {
  Attr* a = find_attr(loc, obj, attr, true);
  if (!a) return -1;
  a->type = 2, a->f = *value;
  return 0;
}
herr_t H5LTget_attribute_string(hid_t loc, const char* obj, const char* attr, char* value) {
  Attr* a = find_attr(loc, obj, attr, false);
  if (!a || a->type != 0) return -1;
  strcpy(value, a->s.c_str());
  return 0;
}
herr_t H5LTget_attribute_long_long(hid_t loc, const char* obj, const char* attr, long long* value) {
  Attr* a = find_attr(loc, obj, attr, false);
  if (!a || a->type != 1) return -1;
  *value = a->i;
  return 0;
}
herr_t H5LTget_attribute_float(hid_t loc, const char* obj, const char* attr, float* value) {
  Attr* a = find_attr(loc, obj, attr, false);
  if (!a || a->type != 2) return -1;
  *value = a->f;
  return 0;
}

}  // extern "C"
