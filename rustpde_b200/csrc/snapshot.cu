// snapshot.cu -- Navier2D::write / read (src/navier/navier.rs:956-1014, src/field/write.rs:82-115,
// src/field/read.rs:56-122): checkpoint / restart with the reference's group / dataset layout.
//
// The reference stores HDF5 (hdf5-interface/src/lib.rs:145-277).  Neither libhdf5 nor h5py exists in this image,
// so the datasets are stored under the SAME names in a flat little-endian container ("RPSNAP1", below);
// rustpde_b200/snapshot.py reads / writes the same container and converts it to / from a real .h5 file wherever
// h5py is installed (`python -m rustpde_b200.snapshot to-h5 in.rpsnap out.h5`).
//
//   magic    8 bytes  "RPSNAP1\0"
//   count    u32      number of datasets
//   dataset  u32 name_len, name (UTF-8, '/'-separated like an HDF5 path, e.g. "temp/vhat_re"),
//            u32 ndim, u64 dims[ndim], f64 data[prod(dims)] (row-major)
//
// Datasets (exactly what write_return_result writes): for g in temp, ux, uy, pres: g/v and g/vhat (real spectral
// space) or g/vhat_re + g/vhat_im (complex); x, dx, y, dy at the root; scalars (ndim = 0) time, ra, pr, nu, kappa.
// As in the reference, temp/v includes the boundary-condition field while temp/vhat does not (navier.rs:988-991).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>

#include "model.h"

namespace rp {

namespace {
struct Dataset {
  std::vector<uint64_t> dims;
  std::vector<double> data;
};
typedef std::map<std::string, Dataset> Snapshot;

void put(Snapshot& s, const std::string& name, std::vector<uint64_t> dims, std::vector<double> data) {
  Dataset d;
  d.dims = std::move(dims);
  d.data = std::move(data);
  s[name] = std::move(d);
}

void write_file(const char* path, const Snapshot& s) {
  FILE* f = fopen(path, "wb");
  if (!f) throw Error(RP_ERR_INVALID, std::string("cannot open for writing: ") + path);
  bool ok = fwrite("RPSNAP1\0", 1, 8, f) == 8;
  const uint32_t cnt = (uint32_t)s.size();
  ok = ok && fwrite(&cnt, 4, 1, f) == 1;
  for (const auto& kv : s) {
    const uint32_t nl = (uint32_t)kv.first.size(), nd = (uint32_t)kv.second.dims.size();
    ok = ok && fwrite(&nl, 4, 1, f) == 1 && fwrite(kv.first.data(), 1, nl, f) == nl && fwrite(&nd, 4, 1, f) == 1;
    if (nd) ok = ok && fwrite(kv.second.dims.data(), 8, nd, f) == nd;
    ok = ok && fwrite(kv.second.data.data(), 8, kv.second.data.size(), f) == kv.second.data.size();
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw Error(RP_ERR_INVALID, std::string("short write: ") + path);
}

Snapshot read_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) throw Error(RP_ERR_INVALID, std::string("cannot open: ") + path);
  Snapshot s;
  char magic[8];
  uint32_t cnt = 0;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "RPSNAP1\0", 8) == 0 && fread(&cnt, 4, 1, f) == 1;
  for (uint32_t i = 0; ok && i < cnt; ++i) {
    uint32_t nl = 0, nd = 0;
    ok = fread(&nl, 4, 1, f) == 1 && nl < 4096;
    std::string name(nl, '\0');
    ok = ok && fread(&name[0], 1, nl, f) == nl && fread(&nd, 4, 1, f) == 1 && nd <= 8;
    Dataset d;
    d.dims.resize(nd);
    if (ok && nd) ok = fread(d.dims.data(), 8, nd, f) == nd;
    uint64_t tot = 1;
    for (uint64_t x : d.dims) tot *= x;
    if (ok) {
      d.data.resize(tot);
      ok = fread(d.data.data(), 8, tot, f) == tot;
    }
    if (ok) s[name] = std::move(d);
  }
  fclose(f);
  if (!ok) throw Error(RP_ERR_INVALID, std::string("not a valid RPSNAP1 file: ") + path);
  return s;
}
}  // namespace

void Navier2D::write_snapshot(const char* path) {
  Snapshot s;
  for (Field2* f : {temp.get(), ux.get(), uy.get(), pres0.get(), field.get()}) f->stream = stream;
  // physical boundary field (navier.rs:988-991): backward of the ortho coefficients through the work field
  std::vector<double> bc((size_t)nx * ny);
  copy_bc_to_field();
  field->backward();
  field->v.download(bc.data(), stream);
  rt::sync(stream);
  const char* names[4] = {"temp", "ux", "uy", "pres"};
  Field2* flds[4] = {temp.get(), ux.get(), uy.get(), pres0.get()};
  for (int i = 0; i < 4; ++i) {
    Field2& f = *flds[i];
    f.backward();
    std::vector<double> v((size_t)f.n0 * f.n1);
    f.v.download(v.data(), stream);
    const size_t w = f.cplx ? 2 : 1;
    std::vector<double> vh((size_t)f.m0 * f.m1 * w);
    f.vhat.download(vh.data(), stream);
    rt::sync(stream);
    if (i == 0)
      for (size_t k = 0; k < v.size(); ++k) v[k] += bc[k];
    const std::string g = names[i];
    put(s, g + "/v", {(uint64_t)f.n0, (uint64_t)f.n1}, std::move(v));
    if (f.cplx) {
      std::vector<double> re((size_t)f.m0 * f.m1), im((size_t)f.m0 * f.m1);
      for (size_t k = 0; k < re.size(); ++k) re[k] = vh[2 * k], im[k] = vh[2 * k + 1];
      put(s, g + "/vhat_re", {(uint64_t)f.m0, (uint64_t)f.m1}, std::move(re));
      put(s, g + "/vhat_im", {(uint64_t)f.m0, (uint64_t)f.m1}, std::move(im));
    } else {
      put(s, g + "/vhat", {(uint64_t)f.m0, (uint64_t)f.m1}, std::move(vh));
    }
    // every field rewrites the root grids (write.rs:103-106); the last one written is pres
    put(s, "x", {(uint64_t)f.x[0].size()}, f.x[0]);
    put(s, "dx", {(uint64_t)f.dx[0].size()}, f.dx[0]);
    put(s, "y", {(uint64_t)f.x[1].size()}, f.x[1]);
    put(s, "dy", {(uint64_t)f.dx[1].size()}, f.dx[1]);
  }
  put(s, "time", {}, {time});
  put(s, "ra", {}, {ra});
  put(s, "pr", {}, {pr});
  put(s, "nu", {}, {nu});
  put(s, "kappa", {}, {ka});
  write_file(path, s);
}

// vhat of temp / ux / uy / pres and the time (navier.rs:963-972).  A stored array of another shape is copied into
// the top-left block both shapes share and the rest of vhat keeps its current values -- exactly what broadcast_2d
// does (read.rs:113-122) -- which is how the reference restarts on a finer or coarser grid.
void Navier2D::read_snapshot(const char* path) {
  const Snapshot s = read_file(path);
  const char* names[4] = {"temp", "ux", "uy", "pres"};
  Field2* flds[4] = {temp.get(), ux.get(), uy.get(), pres0.get()};
  for (int i = 0; i < 4; ++i) {
    Field2& f = *flds[i];
    f.stream = stream;
    const std::string g = names[i];
    const Dataset *re = nullptr, *im = nullptr;
    if (f.cplx) {
      auto a = s.find(g + "/vhat_re"), b = s.find(g + "/vhat_im");
      if (a == s.end() || b == s.end()) {
        fprintf(stderr, "Error while reading file \"%s\".\n", path);  // read.rs: printed and skipped
        continue;
      }
      re = &a->second, im = &b->second;
    } else {
      auto a = s.find(g + "/vhat");
      if (a == s.end()) {
        fprintf(stderr, "Error while reading file \"%s\".\n", path);
        continue;
      }
      re = &a->second;
    }
    if (re->dims.size() != 2 || (im && im->dims != re->dims)) throw Error(RP_ERR_SHAPE, "snapshot: vhat must be 2-D");
    const size_t r0 = (size_t)re->dims[0], c0 = (size_t)re->dims[1];
    const size_t w = f.cplx ? 2 : 1;
    std::vector<double> vh((size_t)f.m0 * f.m1 * w);
    if (r0 != (size_t)f.m0 || c0 != (size_t)f.m1) {
      printf("Attention! Broadcast from shape [%zu, %zu] to shape [%d, %d].\n", r0, c0, f.m0, f.m1);
      f.vhat.download(vh.data(), stream);
      rt::sync(stream);
    }
    const size_t rr = std::min(r0, (size_t)f.m0), cc = std::min(c0, (size_t)f.m1);
    for (size_t a = 0; a < rr; ++a)
      for (size_t b = 0; b < cc; ++b) {
        vh[(a * f.m1 + b) * w] = re->data[a * c0 + b];
        if (im) vh[(a * f.m1 + b) * w + 1] = im->data[a * c0 + b];
      }
    f.vhat.upload(vh.data(), stream);
    ++f.vhat_version;
    f.backward();
    rt::sync(stream);
  }
  auto t = s.find("time");
  if (t == s.end() || t->second.data.empty()) throw Error(RP_ERR_INVALID, "snapshot: no `time` scalar");
  time = t->second.data[0];
}

}  // namespace rp
