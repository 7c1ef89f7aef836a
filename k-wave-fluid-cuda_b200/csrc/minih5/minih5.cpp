// minih5.cpp -- implementation of the HDF5 API subset declared in hdf5.h / hdf5_hl.h over an in-memory object tree.
//
// KWH5 container (little endian), written on H5Fclose of a writable file:
//   magic   "KWH5\x00\x01\x00\x00"
//   u64     number of records
//   record  u32 path_len, path (absolute, "/" = root group)
//           u8  kind (0 group, 1 float32 dataset, 2 uint64 dataset)
//           u32 nattrs, then per attribute: u32 name_len, name, u8 type (0 string, 1 int64, 2 float32),
//               payload (string: u32 len + bytes | int64 | float32)
//           datasets only: u32 rank, u64 dims[rank], u32 chunk_rank, u64 chunk[chunk_rank], u32 deflate_level, raw data
// Records appear parent-before-child.  tools/kwh5.py reads and writes the same format from Python.
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

struct Node {
  bool is_group = true;
  int dtype = 1;  // 1 float32, 2 uint64
  std::vector<hsize_t> dims, chunk;
  unsigned deflate = 0;
  std::vector<uint8_t> data;
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

uint64_t count_nodes(const Node& n) {
  uint64_t c = 1;
  for (auto& kv : n.children) c += count_nodes(*kv.second);
  return c;
}
void write_node(FILE* f, const Node& n, const std::string& path) {
  put_str(f, path);
  put<uint8_t>(f, n.is_group ? 0 : (uint8_t)n.dtype);
  put<uint32_t>(f, (uint32_t)n.attrs.size());
  for (auto& kv : n.attrs) {
    put_str(f, kv.first);
    put<uint8_t>(f, (uint8_t)kv.second.type);
    if (kv.second.type == 0) put_str(f, kv.second.s);
    else if (kv.second.type == 1) put<int64_t>(f, kv.second.i);
    else put<float>(f, kv.second.f);
  }
  if (!n.is_group) {
    put<uint32_t>(f, (uint32_t)n.dims.size());
    for (auto d : n.dims) put<uint64_t>(f, d);
    put<uint32_t>(f, (uint32_t)n.chunk.size());
    for (auto d : n.chunk) put<uint64_t>(f, d);
    put<uint32_t>(f, n.deflate);
    fwrite(n.data.data(), 1, n.data.size(), f);
  }
  for (auto& name : n.order) write_node(f, *n.children.at(name), join(path, name));
}
bool save(const File& file) {
  FILE* f = fopen(file.path.c_str(), "wb");
  if (!f) return false;
  fwrite(kMagic, 1, 8, f);
  put<uint64_t>(f, count_nodes(file.root));
  write_node(f, file.root, "/");
  const bool ok = fflush(f) == 0;
  fclose(f);
  return ok;
}
bool load(File* file) {
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
  while (fi < fr.size() && mi < mr.size()) {
    const size_t n = std::min(fr[fi].second - fo, mr[mi].second - mo);
    uint8_t* fp = ds->data.data() + (fr[fi].first + fo) * es;
    uint8_t* mp = mem + (mr[mi].first + mo) * es;
    if ((fr[fi].first + fo + n) * es > ds->data.size()) return -1;
    if (to_file) memcpy(fp, mp, n * es);
    else memcpy(mp, fp, n * es);
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
  const bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0;
  fclose(f);
  return ok ? 1 : 0;
}
herr_t H5Fget_filesize(hid_t file, hsize_t* size) {
  Handle* h = get(file);
  if (!h || h->kind != kFile) return -1;
  // size the file would have on disk now
  struct Acc {
    static hsize_t of(const Node& n) {
      hsize_t s = 64 + n.data.size();
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
  if (Handle* pl = dcpl ? get(dcpl) : nullptr) n->chunk = pl->plist.chunk, n->deflate = pl->plist.deflate;
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
