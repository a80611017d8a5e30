// ConvertToBinary — drop-in for source/mesh_stream/ConvertToBinary.cpp on B200.
//
// Geometry: disparity PFM -> camera mesh (.vtx float32 xyz, .idx uint32 x 3, optional .obj) through libderp_b200.so
// (derp_camera_mesh_simplified: mesh_util::getVertexesEquiError / getFaces / applyMaskToVertexesAndFaces on the GPU, then
// render::MeshSimplifier's contraction sweeps down to --triangles, sequential host code like the reference's), the rescaled
// `<rig>_fused.json`, and the striped fusion of the produced files (BinaryFusionUtil.h).
// Colour: .bc7 (the reference's default; derp_bc7_compress_image: conversion, gamma correction, packing and the block encoder
// of bc7_util::compressBC7 in one CUDA kernel) and the uncompressed .rgba stream (host bytes).
// The rasterised pfm format (mesh_util::writePfm) is host code like the reference's (io.h rasterMesh; its 2 x 2 solves are by
// elimination, not Eigen's QR: pixel centres exactly on an edge can fall the other way, values agree to float rounding).
// --color_scale < 1 (the UI's export below full width): conversion, cv::resize INTER_AREA and gamma are host stages pinned to
// cv2 (area_resize.h), the encoder takes the packed RGBA8 surface.
#include <set>
#include <thread>

#include "io.h"

const std::string kUsage = R"(
       - Expects all files to be in the format <dir>/<camera>/<frame>.extension

       If <color> is specified:
       - Read .png files and save them as .rgba files in <bin> folder
       If <disparity> is specified:
       - Read .pfm files and save them as .vtx and .idx files in <bin> folder

       <bin> folder is created for each frame if it does not exist

       If <rgba> is specified:
       - Convert color image into an RGBA binary stream

       If <obj> is specified:
       - Read .vtx and .idx files from <bin> and save .obj files to <obj> folder

       - Example:
         ./ConvertToBinary \
         --color=/path/to/video/color \
         --rig=/path/to/rigs/rig.json \
         --first=000000 \
         --last=000000 \
         --disparity=/path/to/output/disparity \
         --bin=/path/to/output/bin \
         --fused=/path/to/output/fused
     )";

DEFINE_string(bin, "bin", "output directory containing binary data");
DEFINE_string(cameras, "", "cameras to render (comma-separated)");
DEFINE_string(color, "", "path to input color images");
DEFINE_double(color_scale, 1, "optional color scale before compression & fusion (>= 1 = no scale)");
DEFINE_double(depth_scale, 1, "optional depthmap scale before simplification (>= 1 = no scale)");
DEFINE_string(disparity, "", "path to disparity images (pfm)");
DEFINE_string(first, "", "first frame to process (lexical) (required)");
DEFINE_string(foreground_masks, "", "path to foreground masks specifying regions to include in per-frame geometry");
DEFINE_int32(fuse_strip, 1, "number of strip files");
DEFINE_string(fused, "", "output directory containing fused binary data, ready for playback");
DEFINE_double(gamma_correction, 2.2 / 1.8, "exponent to raise color channels before BC7 encoding");
DEFINE_string(last, "", "last frame to process (lexical) (required)");
DEFINE_string(output_formats, "idx,vtx,bc7", "saved formats, comma separated (idx, vtx, bc7 default; rgba, pfm, obj also supported)");
DEFINE_string(rig, "", "path to camera rig .json (required)");
DEFINE_bool(run_conversion, true, "whether or not to run binary conversion");
DEFINE_double(tear_ratio, 0.95, "depth ratio that causes mesh to tear");
DEFINE_int32(threads, -1, "number of threads (-1 = max allowed, 0 = no threading)");
DEFINE_int32(triangles, 150000, "number of triangles per camera mesh (<= 0: no simplification)");
DEFINE_int32(gpu, 0, "first CUDA device to use");
DEFINE_int32(gpus, 1, "number of GPUs of this box to shard (frame, camera) meshes across");

#define DERP_CALL(expr)                                                 \
  do {                                                                  \
    const int rc_ = (expr);                                             \
    if (rc_ != 0) LOG(FATAL) << #expr << " failed: " << derp_last_error(); \
  } while (0)

static const uint64_t kStripeSize = 512 * 1024;  // StripedFile.h:19

static bool contains(const std::vector<std::string>& v, const std::string& s) { return std::find(v.begin(), v.end(), s) != v.end(); }

// width / height of a PNG from its IHDR chunk
static void pngSize(const fs::path& p, int* w, int* h) {
  std::ifstream f(p, std::ios::binary);
  uint8_t hdr[24];
  f.read(reinterpret_cast<char*>(hdr), 24);
  CHECK(f.gcount() == 24 && std::memcmp(hdr + 12, "IHDR", 4) == 0) << "not a PNG: " << p.string();
  *w = (int)io::be32(hdr + 16);
  *h = (int)io::be32(hdr + 20);
}

// folly::json with sort_keys, pretty_formatting, DtoaMode::FIXED and `digits` decimals (Camera.cpp:305-312)
static void writeJson(std::ostream& os, const io::Json& j, int digits, int indent = 0) {
  const std::string pad((size_t)indent * 2, ' '), pad2((size_t)indent * 2 + 2, ' ');
  switch (j.type) {
    case io::Json::Null: os << "null"; break;
    case io::Json::Bool: os << (j.b ? "true" : "false"); break;
    case io::Json::Num: {
      char buf[64];
      if (j.num == std::floor(j.num) && std::fabs(j.num) < 1e15 && digits < 0) snprintf(buf, sizeof(buf), "%lld", (long long)j.num);
      else snprintf(buf, sizeof(buf), "%.*f", digits < 0 ? 10 : digits, j.num);
      os << buf;
      break;
    }
    case io::Json::Str: os << '"' << j.str << '"'; break;
    case io::Json::Arr:
      os << "[";
      for (size_t i = 0; i < j.arr.size(); ++i) {
        os << (i ? ",\n" : "\n") << pad2;
        writeJson(os, j.arr[i], digits, indent + 1);
      }
      os << (j.arr.empty() ? "" : "\n" + pad) << "]";
      break;
    case io::Json::Obj: {
      std::vector<std::pair<std::string, io::Json>> kv = j.obj;
      std::sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      os << "{";
      for (size_t i = 0; i < kv.size(); ++i) {
        os << (i ? ",\n" : "\n") << pad2 << '"' << kv[i].first << "\": ";
        writeJson(os, kv[i].second, digits, indent + 1);
      }
      os << (kv.empty() ? "" : "\n" + pad) << "}";
      break;
    }
  }
}
static io::Json jnum(double v) {
  io::Json j;
  j.type = io::Json::Num;
  j.num = v;
  return j;
}
static io::Json jstr(const std::string& s) {
  io::Json j;
  j.type = io::Json::Str;
  j.str = s;
  return j;
}
static io::Json jvec(const double* v, int n) {
  io::Json j;
  j.type = io::Json::Arr;
  for (int i = 0; i < n; ++i) j.arr.push_back(jnum(v[i]));
  return j;
}

// Camera::serialize (Camera.cpp:158-177) of the (possibly rescaled) camera.  The reference writes the re-unitarised
// rotation rows; this writes the rows of the input rig, which describe the same camera.
static io::Json serializeCamera(const DerpCameraDesc& c, const std::string& id) {
  static const char* names[] = {"FTHETA", "RECTILINEAR", "EQUISOLID", "ORTHOGRAPHIC"};
  io::Json j;
  j.type = io::Json::Obj;
  io::Json version = jnum(1);
  j.obj = {{"version", version},          {"type", jstr(names[c.type])},     {"origin", jvec(c.origin, 3)},
           {"forward", jvec(c.forward, 3)}, {"up", jvec(c.up, 3)},             {"right", jvec(c.right, 3)},
           {"resolution", jvec(c.resolution, 2)}, {"focal", jvec(c.focal, 2)}, {"id", jstr(id)}};
  if (c.has_principal && (c.principal[0] != c.resolution[0] / 2 || c.principal[1] != c.resolution[1] / 2))
    j.obj.push_back({"principal", jvec(c.principal, 2)});
  if (c.distortion[0] != 0 || c.distortion[1] != 0 || c.distortion[2] != 0) j.obj.push_back({"distortion", jvec(c.distortion, 3)});
  if (c.has_fov) j.obj.push_back({"fov", jnum(c.fov)});
  return j;
}

// mesh_util::writeObj (MeshUtil.h:95-133) from the .vtx / .idx just written (ConvertToBinary.cpp:226-232)
static void writeObj(const std::vector<float>& v, const std::vector<uint32_t>& f, const fs::path& path) {
  FILE* fp = fopen(path.c_str(), "w");
  CHECK(fp) << "file open failed: " << path.string();
  for (size_t i = 0; i + 2 < v.size(); i += 3) fprintf(fp, "v %g %g %g\n", (double)v[i], (double)v[i + 1], (double)v[i + 2]);
  for (size_t i = 0; i + 2 < f.size(); i += 3) fprintf(fp, "f %d %d %d\n", (int)f[i] + 1, (int)f[i + 1] + 1, (int)f[i + 2] + 1);
  fclose(fp);
}

// ---- fusion (BinaryFusionUtil.h:26-84, StripedFile.h:96-101) -----------------------------------------------------
static uint64_t alignUp(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
static void calcStripe(uint64_t& local, uint64_t& disk, uint64_t global, uint64_t diskCount) {
  const uint64_t stripe = global / kStripeSize;
  local = (stripe / diskCount) * kStripeSize;
  disk = stripe % diskCount;
}
static void addFile(std::vector<FILE*>& disks, uint64_t& offset, const fs::path& filename) {
  const uint64_t aligned = alignUp(offset, kStripeSize);
  uint64_t end = offset == aligned ? offset + kStripeSize : aligned;
  uint64_t size = fs::file_size(filename);
  FILE* file = fopen(filename.c_str(), "rb");
  CHECK(file) << "cannot open " << filename.string();
  LOG(INFO) << "Fusing " << filename.string() << "...";
  while (size) {
    std::vector<uint8_t> buffer(std::min(size, end - offset));
    CHECK_EQ(fread(buffer.data(), 1, buffer.size(), file), buffer.size()) << "Error reading buffer data";
    uint64_t local, disk;
    calcStripe(local, disk, offset, disks.size());
    fwrite(buffer.data(), 1, buffer.size(), disks[disk]);
    offset += buffer.size();
    end = offset + kStripeSize;
    size -= buffer.size();
  }
  fclose(file);
}
static void padStripe(std::vector<FILE*>& disks, uint64_t& offset) {
  const uint64_t aligned = alignUp(offset, kStripeSize);
  if (offset == aligned) return;
  std::vector<uint8_t> buffer(aligned - offset, 0x5A);
  uint64_t local, disk;
  calcStripe(local, disk, offset, disks.size());
  fwrite(buffer.data(), 1, buffer.size(), disks[disk]);
  offset += buffer.size();
}
static io::Json jobj() {
  io::Json j;
  j.type = io::Json::Obj;
  return j;
}

static void fuse(const io::Rig& rig, const std::vector<int>& cams, const std::vector<std::string>& formats, const fs::path& rigOut) {
  std::vector<FILE*> disks;
  fs::create_directories(FLAGS_fused);
  for (int i = 0; i < FLAGS_fuse_strip; ++i) {
    const std::string diskName = FLAGS_fused + "/fused_" + std::to_string(i) + ".bin";
    FILE* disk = fopen(diskName.c_str(), "wb");
    CHECK(disk) << "Failed to open " << diskName;
    disks.push_back(disk);
  }
  uint64_t offset = 0;
  io::Json frames = jobj();
  const int first = std::stoi(FLAGS_first), numFrames = std::stoi(FLAGS_last) - first + 1;
  for (int iFrame = 0; iFrame < numFrames; ++iFrame) {
    const std::string frameName = io::zeroPad(iFrame + first);
    LOG(INFO) << "Fusing frame " << frameName << "...";
    io::Json frame = jobj();
    for (int c : cams) {
      const uint64_t begin = offset;
      io::Json camera = jobj();
      for (const std::string& format : formats) {
        if (format.empty()) continue;
        const uint64_t b = offset;
        addFile(disks, offset, fs::path(FLAGS_bin) / rig.ids[c] / (frameName + "." + format));
        io::Json e = jobj();
        e.obj = {{"offset", jnum((double)b)}, {"size", jnum((double)(offset - b))}};
        camera.obj.push_back({"." + format, e});
      }
      camera.obj.push_back({"offset", jnum((double)begin)});
      camera.obj.push_back({"size", jnum((double)(offset - begin))});
      padStripe(disks, offset);
      frame.obj.push_back({rig.ids[c], camera});
    }
    frames.obj.push_back({frameName, frame});
  }
  io::Json meta = jobj(), little;
  little.type = io::Json::Bool;
  little.b = true;
  meta.obj = {{"isLittleEndian", little}};
  io::Json catalog = jobj();
  catalog.obj = {{"metadata", meta}, {"frames", frames}};
  std::ofstream os(FLAGS_fused + "/fused.json", std::ios::binary);
  writeJson(os, catalog, -1);
  os << "\n";
  for (FILE* disk : disks) fclose(disk);
  fs::copy_file(rigOut, fs::path(FLAGS_fused) / rigOut.filename(), fs::copy_options::overwrite_existing);
}

int main(int argc, char** argv) {
  flags::initDep(argc, argv, kUsage);
  CHECK_LE(FLAGS_color_scale, 1.);
  CHECK_LE(FLAGS_depth_scale, 1.);
  CHECK_NE(FLAGS_rig, "");
  CHECK_NE(FLAGS_first, "");
  CHECK_NE(FLAGS_last, "");
  io::Rig rig = io::loadRig(FLAGS_rig);
  const std::vector<int> cams = io::filterDestinations(rig, FLAGS_cameras);
  CHECK_GT(cams.size(), 0u) << "No cameras to convert";

  std::vector<std::string> formats;
  {
    std::stringstream ss(FLAGS_output_formats);
    for (std::string f; std::getline(ss, f, ',');) formats.push_back(f);
  }
  const std::set<std::string> supported = {"idx", "vtx", "bc7", "obj", "pfm", "rgba"};
  for (const std::string& f : formats) CHECK(f.empty() || supported.count(f)) << "Invalid output format specified: " << f;
  const int firstFrame = std::stoi(FLAGS_first), numFrames = std::stoi(FLAGS_last) - firstFrame + 1;
  CHECK_GT(numFrames, 0);
  const bool wantRgba = !FLAGS_color.empty() && contains(formats, "rgba");
  const bool wantBc7 = !FLAGS_color.empty() && contains(formats, "bc7");
  const bool wantDepth = !FLAGS_disparity.empty() &&
      (contains(formats, "idx") || contains(formats, "vtx") || contains(formats, "obj") || contains(formats, "pfm"));

  // resizeRig (ConvertToBinary.cpp:322-343): camera resolutions follow the (scaled) colour images
  if (!FLAGS_color.empty()) {
    for (int c : cams) {
      int w, h;
      pngSize(io::imagePath(FLAGS_color, rig.ids[c], FLAGS_first), &w, &h);
      if (FLAGS_color_scale < 1) io::scaledSize(w, h, FLAGS_color_scale, &w, &h);  // cv_util::scaleImage
      DerpCameraDesc& cam = rig.cams[c];
      const float xScale = float(w) / cam.resolution[0], yScale = float(h) / cam.resolution[1];
      CHECK_EQ(xScale, yScale) << "Aspect ratio must be kept. " << cam.resolution[0] << "x" << cam.resolution[1] << " vs " << w << "x" << h;
      if (xScale != 1) {  // Camera::rescale (Camera.cpp:217-223)
        if (!cam.has_principal) {
          cam.principal[0] = cam.resolution[0] / 2;
          cam.principal[1] = cam.resolution[1] / 2;
          cam.has_principal = 1;
        }
        for (int k = 0; k < 2; ++k) {
          const double newRes = xScale * cam.resolution[k];
          cam.principal[k] *= newRes / cam.resolution[k];
          cam.focal[k] *= newRes / cam.resolution[k];
          cam.resolution[k] = newRes;
        }
      }
    }
  }

  const fs::path rigOut = fs::path(FLAGS_bin) / (fs::path(FLAGS_rig).stem().string() + "_fused.json");
  if (FLAGS_run_conversion) {
    if (wantDepth)
      for (int c : cams)
        for (int f = 0; f < numFrames; ++f) {
          CHECK(fs::is_regular_file(io::imagePath(FLAGS_disparity, rig.ids[c], io::zeroPad(firstFrame + f), ".pfm")))
              << "Missing disparity for camera " << rig.ids[c] << " frame " << io::zeroPad(firstFrame + f);
          CHECK_EQ(rig.cams[c].focal[0], -rig.cams[c].focal[1]) << "pixels are not square";  // Camera::getScalarFocal
        }
    struct Task {
      int cam, frame;
    };
    std::vector<Task> tasks;
    for (int f = 0; f < numFrames; ++f)
      for (int c : cams) tasks.push_back(Task{c, firstFrame + f});
    const int G = std::max(1, std::min<int>(FLAGS_gpus, (int)tasks.size()));
    // --threads like ThreadPool.h:30-45 (-1 = all cores, 0 = inline): one (frame, camera) conversion per task as in the
    // reference (ConvertToBinary.cpp:361-368).  The contraction sweeps of the simplifier are host work, so the workers
    // are host threads; worker k sends its meshes to GPU k mod --gpus.
    const int want = FLAGS_threads < 0 ? (int)std::max(1u, std::thread::hardware_concurrency()) : std::max(1, FLAGS_threads);
    const int workers = std::max(G, std::min<int>(std::min(want, 64), (int)tasks.size()));
    LOG(INFO) << "backend " << derp_backend() << ", " << G << " GPU(s), " << workers << " host worker(s)";
    std::vector<std::thread> threads;
    for (int k = 0; k < workers && wantDepth; ++k)
      threads.emplace_back([&, k] {
        const int device = FLAGS_gpu + k % G;
        for (size_t t = k; t < tasks.size(); t += workers) {
          const std::string &id = rig.ids[tasks[t].cam], frame = io::zeroPad(tasks[t].frame);
          LOG(INFO) << "Converting depth: frame " << frame << ", camera " << id << "...";
          int w, h, mw = 0, mh = 0, gw, gh;
          const std::vector<float> disparity = io::readPfm(io::imagePath(FLAGS_disparity, id, frame, ".pfm"), &w, &h);
          std::vector<uint8_t> mask;
          if (!FLAGS_foreground_masks.empty()) {
            mask = io::loadMask(io::imagePath(FLAGS_foreground_masks, id, frame), &mw, &mh);  // loadImage<bool>: 0 / 1
          }
          DERP_CALL(derp_camera_mesh_size(w, h, FLAGS_depth_scale, &gw, &gh));
          std::vector<float> vtx((size_t)gw * gh * 3);
          std::vector<uint32_t> idx((size_t)gw * gh * 6);
          uint64_t nv = 0, nf = 0;
          const DerpCameraDesc& cam = rig.cams[tasks[t].cam];
          if (FLAGS_triangles > 0) LOG(INFO) << "Target number of faces: " << FLAGS_triangles;
          DERP_CALL(derp_camera_mesh_simplified(device, disparity.data(), w, h, FLAGS_depth_scale, cam.resolution[0],
                                                cam.resolution[1], cam.focal[0], (float)FLAGS_tear_ratio,
                                                mask.empty() ? nullptr : mask.data(), mw, mh, FLAGS_triangles, vtx.data(),
                                                idx.data(), &nv, &nf));
          vtx.resize(nv * 3);
          idx.resize(nf * 3);
          LOG(INFO) << "camera " << id << ": " << nv << " vertexes, " << nf << " faces";
          const fs::path fnVtx = io::imagePath(FLAGS_bin, id, frame, ".vtx"), fnIdx = io::imagePath(FLAGS_bin, id, frame, ".idx");
          fs::create_directories(fnVtx.parent_path());
          if (contains(formats, "idx") || contains(formats, "vtx")) {  // mesh_util::writeDepth writes both
            std::ofstream(fnVtx, std::ios::binary).write(reinterpret_cast<const char*>(vtx.data()), vtx.size() * sizeof(float));
            std::ofstream(fnIdx, std::ios::binary).write(reinterpret_cast<const char*>(idx.data()), idx.size() * sizeof(uint32_t));
          }
          if (contains(formats, "pfm")) {  // mesh_util::writePfm(depth, cam.resolution, vertexes, faces, ...): the mesh rasterised
            const std::vector<float> raster =  // back onto the (possibly --depth_scale'd) depth grid; host code like the reference's
                io::rasterMesh(vtx.data(), idx.data(), nf, gw, gh, cam.resolution[0], cam.resolution[1]);
            io::writePfm(io::imagePath(FLAGS_bin, id, frame, ".pfm"), raster.data(), gw, gh);
          }
          if (contains(formats, "obj")) writeObj(vtx, idx, io::imagePath(FLAGS_bin, id, frame, ".obj"));
        }
      });
    for (auto& t : threads) t.join();

    // convertColor, bc7 (ConvertToBinary.cpp:131-137): bc7_util::compressBC7(image, path, --gamma_correction, no DDS header).
    // The image goes to the GPU as stored (B, G, R, 8 or 16 bits); conversion to float, gamma correction, RGBA packing and the
    // block encoder are one kernel (derp_bc7_compress_image).  PNG decoding is the host cost: one worker per image.
    if (wantBc7) {
      threads.clear();
      for (int k = 0; k < workers; ++k)
        threads.emplace_back([&, k] {
          const int device = FLAGS_gpu + k % G;
          for (size_t t = k; t < tasks.size(); t += workers) {
            const std::string &id = rig.ids[tasks[t].cam], frame = io::zeroPad(tasks[t].frame);
            LOG(INFO) << "Converting color: frame " << frame << ", camera " << id << "...";
            const io::Image img = io::loadUnchanged(io::imagePath(FLAGS_color, id, frame));
            if (FLAGS_color_scale < 1) {
              // loadScaledImage<Vec4f>(..., --color_scale, INTER_AREA): conversion, area resize and gamma are host stages
              // (the reference's arithmetic incl. its powf); the block encoder takes the packed surface
              int sw, sh;
              const std::vector<uint8_t> rgba = io::bc7SurfaceScaled(img, FLAGS_color_scale, (float)FLAGS_gamma_correction, &sw, &sh);
              std::vector<uint8_t> blocks((size_t)sw * sh);
              DERP_CALL(derp_bc7_compress(device, rgba.data(), sw, sh, blocks.data()));
              const fs::path out = io::imagePath(FLAGS_bin, id, frame, ".bc7");
              fs::create_directories(out.parent_path());
              std::ofstream(out, std::ios::binary).write(reinterpret_cast<const char*>(blocks.data()), blocks.size());
              continue;
            }
            CHECK((img.bits == 8 || img.bits == 16) && (img.channels == 1 || img.channels == 3 || img.channels == 4))
                << "Conversion from " << img.channels << " channels to 4 channels not supported";  // CvUtil.h:261-262
            const size_t n = (size_t)img.w * img.h;
            std::vector<uint8_t> pixels(n * 3 * (img.bits / 8));  // B, G, R as cv::cvtColor(GRAY2BGRA / BGRA2BGR) leave them
            for (size_t i = 0; i < n; ++i)
              for (int c = 0; c < 3; ++c) {
                const uint16_t v = img.u[i * img.channels + (img.channels == 1 ? 0 : c)];
                if (img.bits == 8)
                  pixels[i * 3 + c] = (uint8_t)v;
                else
                  std::memcpy(&pixels[(i * 3 + c) * 2], &v, 2);
              }
            std::vector<uint8_t> blocks(n);
            DERP_CALL(derp_bc7_compress_image(device, pixels.data(), img.bits, 3, img.w, img.h, (float)FLAGS_gamma_correction,
                                              blocks.data()));
            const fs::path out = io::imagePath(FLAGS_bin, id, frame, ".bc7");
            fs::create_directories(out.parent_path());
            std::ofstream(out, std::ios::binary).write(reinterpret_cast<const char*>(blocks.data()), blocks.size());
          }
        });
      for (auto& t : threads) t.join();
    }

    if (wantRgba)  // convertColor (ConvertToBinary.cpp:122-147): ".rgba is just uncompressed 8-bit color"
      for (const Task& t : tasks) {
        const std::string &id = rig.ids[t.cam], frame = io::zeroPad(t.frame);
        LOG(INFO) << "Converting color: frame " << frame << ", camera " << id << "...";
        int w, h;
        std::vector<uint8_t> rgba = io::loadRgba8(io::imagePath(FLAGS_color, id, frame), &w, &h);
        if (FLAGS_color_scale < 1) rgba = io::scaleRgba8(rgba, &w, &h, FLAGS_color_scale);
        const fs::path out = io::imagePath(FLAGS_bin, id, frame, ".rgba");
        fs::create_directories(out.parent_path());
        std::ofstream(out, std::ios::binary).write(reinterpret_cast<const char*>(rgba.data()), rgba.size());
      }

    fs::create_directories(FLAGS_bin);
    io::Json root = jobj(), arr;
    arr.type = io::Json::Arr;
    for (int c : cams) arr.arr.push_back(serializeCamera(rig.cams[c], rig.ids[c]));
    root.obj = {{"cameras", arr}};
    std::ofstream os(rigOut, std::ios::binary);
    writeJson(os, root, 10);
    os << "\n";
  }

  if (!FLAGS_fused.empty()) {
    std::vector<std::string> fuseFormats;
    for (const std::string& f : formats)
      if (!f.empty()) fuseFormats.push_back(f);
    fuse(rig, cams, fuseFormats, rigOut);
  }
  return EXIT_SUCCESS;
}
