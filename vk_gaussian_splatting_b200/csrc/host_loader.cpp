// host_loader.cpp — scene files -> SplatSet layout (host C++).
//
// Replaces PlyLoaderAsync::innerLoad (src/ply_loader_async.cpp:291-453) for the three formats the
// reference accepts, with the same post-processing, so the arrays are bit-identical to what the
// reference hands to SplatSetVk:
//   .ply   INRIA 3DGS point cloud, parsed by property NAME (the reference uses miniply,
//          3rdparty/miniply): x y z | opacity | scale_0..2 | rot_0..3 | f_dc_0..2 | f_rest_0..44 (all 45
//          or SH degree 0, :383-395); ascii, binary_little_endian and binary_big_endian;
//          then RDF -> RUB (src/splat_set.h:78-114).
//   .splat antimatter15 32-byte records (:43-183): log(scale), (c/255-0.5)/C0, logit(clamped alpha),
//          (b-128)/128 quaternion bytes in stored order; then RDF -> RUB.
//   .spz   Niantic packed gaussians (3rdparty/spz/src/cc/load-spz.cc:476-596): gzip, 16-byte header,
//          24-bit fixed-point positions, u8 alphas/colours/scales, smallest-three (v3) or first-three
//          (v2) quaternions, u8 SH; unpacked in RUB, quaternion xyzw -> wxyz and SH RGB-inner ->
//          channel-major (src/ply_loader_async.cpp:304-346).
// zlib (system library) is used for the gzip container only.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <functional>
#include <thread>

#include "vkgs_b200.h"

struct vkgs_scene
{
  std::vector<float> positions, f_dc, f_rest, opacity, scale, rotation;
  uint32_t           fRestPerSplat = 0;
  std::string        error;
};

namespace {

thread_local std::string g_loaderError;

bool failLoad(const std::string& msg)
{
  g_loaderError = msg;
  return false;
}

std::string lowerExt(const std::string& path)
{
  const size_t dot = path.find_last_of('.');
  if(dot == std::string::npos)
    return "";
  std::string e = path.substr(dot);
  std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
  return e;
}

bool readFile(const std::string& path, std::vector<uint8_t>& out)
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if(!f.is_open())
    return failLoad("cannot open " + path);
  const std::streamsize n = f.tellg();
  f.seekg(0, std::ios::beg);
  out.resize(static_cast<size_t>(n));
  if(n > 0)
    f.read(reinterpret_cast<char*>(out.data()), n);
  return f ? true : failLoad("read error on " + path);
}

// Read-only view of a whole file: mapped when possible (no copy, no zero fill of a staging vector), read() otherwise.
class FileBytes
{
public:
  FileBytes() = default;
  FileBytes(const FileBytes&)            = delete;
  FileBytes& operator=(const FileBytes&) = delete;
  ~FileBytes()
  {
    if(m_mapped)
      munmap(const_cast<uint8_t*>(m_data), m_size);
  }
  bool open(const std::string& path)
  {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if(fd < 0)
      return failLoad("cannot open " + path);
    struct stat st;
    if(fstat(fd, &st) != 0 || !S_ISREG(st.st_mode))
    {
      ::close(fd);
      return readFile(path, m_fallback) && adoptFallback();
    }
    m_size = static_cast<size_t>(st.st_size);
    if(m_size == 0)
    {
      ::close(fd);
      m_data = nullptr;
      return true;
    }
    void* p = mmap(nullptr, m_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if(p == MAP_FAILED)
      return readFile(path, m_fallback) && adoptFallback();
    madvise(p, m_size, MADV_SEQUENTIAL);
    m_data   = static_cast<const uint8_t*>(p);
    m_mapped = true;
    return true;
  }
  const uint8_t* data() const { return m_data; }
  size_t         size() const { return m_size; }
  uint8_t        operator[](size_t i) const { return m_data[i]; }

private:
  bool adoptFallback()
  {
    m_data = m_fallback.data(), m_size = m_fallback.size();
    return true;
  }
  const uint8_t*       m_data   = nullptr;
  size_t               m_size   = 0;
  bool                 m_mapped = false;
  std::vector<uint8_t> m_fallback;
};

// fn(begin, end) over [0, n) on the host's threads (rows of a scene are independent)
void parallelFor(size_t n, const std::function<void(size_t, size_t)>& fn)
{
  const size_t hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const size_t T  = n < (1u << 16) ? 1 : hw;
  if(T == 1)
  {
    fn(0, n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t             chunk = (n + T - 1) / T;
  for(size_t t = 0; t < T; t++)
  {
    const size_t b = t * chunk, e = std::min(n, b + chunk);
    if(b < e)
      pool.emplace_back([&fn, b, e] { fn(b, e); });
  }
  for(std::thread& th : pool)
    th.join();
}

// SplatSet::convertCoordinates(RDF, RUB): flipP = (1,-1,-1), flipQ = (1,-1,-1) on (x,y,z), per-coefficient
// SH signs (spz coordinateConverter with x match, y and z flipped).
void rdfToRub(vkgs_scene& s)
{
  const float  flipP[3]   = {1.0f, -1.0f, -1.0f};
  const float  flipQ[3]   = {1.0f, -1.0f, -1.0f};
  const float  x = 1.0f, y = -1.0f, z = -1.0f;
  const float  flipSh[15] = {y, z, x, x * y, y * z, 1.0f, x * z, 1.0f, y, x * y * z, y, z, x, z, x};
  const size_t n          = s.positions.size() / 3;
  if(n == 0)
    return;
  const size_t perPoint = s.f_rest.size() / 3 / n;
  const size_t nRot     = s.rotation.size() / 4;
  parallelFor(n, [&](size_t b, size_t e) {
    for(size_t i = b; i < e; ++i)
    {
      s.positions[3 * i + 0] *= flipP[0];
      s.positions[3 * i + 1] *= flipP[1];
      s.positions[3 * i + 2] *= flipP[2];
      if(i < nRot)
      {
        s.rotation[4 * i + 1] *= flipQ[0];
        s.rotation[4 * i + 2] *= flipQ[1];
        s.rotation[4 * i + 3] *= flipQ[2];
      }
      const size_t idx = i * 3 * perPoint;
      for(size_t j = 0; j < perPoint && j < 15; ++j)
      {
        const float flip = flipSh[j];
        s.f_rest[idx + j] *= flip;
        s.f_rest[idx + perPoint + j] *= flip;
        s.f_rest[idx + perPoint * 2 + j] *= flip;
      }
    }
  });
}

// ---------------------------------------------------------------------------------------------
// .ply

enum class PlyType
{
  Char,
  UChar,
  Short,
  UShort,
  Int,
  UInt,
  Float,
  Double,
  None
};

PlyType parseType(const std::string& t)
{
  if(t == "char" || t == "int8")
    return PlyType::Char;
  if(t == "uchar" || t == "uint8")
    return PlyType::UChar;
  if(t == "short" || t == "int16")
    return PlyType::Short;
  if(t == "ushort" || t == "uint16")
    return PlyType::UShort;
  if(t == "int" || t == "int32")
    return PlyType::Int;
  if(t == "uint" || t == "uint32")
    return PlyType::UInt;
  if(t == "float" || t == "float32")
    return PlyType::Float;
  if(t == "double" || t == "float64")
    return PlyType::Double;
  return PlyType::None;
}

size_t typeSize(PlyType t)
{
  switch(t)
  {
    case PlyType::Char:
    case PlyType::UChar:
      return 1;
    case PlyType::Short:
    case PlyType::UShort:
      return 2;
    case PlyType::Int:
    case PlyType::UInt:
    case PlyType::Float:
      return 4;
    case PlyType::Double:
      return 8;
    default:
      return 0;
  }
}

struct PlyProp
{
  std::string name;
  PlyType     type      = PlyType::None;
  PlyType     countType = PlyType::None;  // != None for list properties
  size_t      offset    = 0;              // byte offset inside a fixed-size row
};

struct PlyElement
{
  std::string          name;
  size_t               count = 0;
  std::vector<PlyProp> props;
  bool                 fixedSize = true;
  size_t               rowBytes  = 0;
};

template <typename T>
T loadScalar(const uint8_t* p, bool swap)
{
  uint8_t b[sizeof(T)];
  if(swap)
    for(size_t i = 0; i < sizeof(T); i++)
      b[i] = p[sizeof(T) - 1 - i];
  else
    std::memcpy(b, p, sizeof(T));
  T v;
  std::memcpy(&v, b, sizeof(T));
  return v;
}

// value of a binary scalar converted to float the way miniply's extract_properties(..., Float) does
float scalarToFloat(const uint8_t* p, PlyType t, bool swap)
{
  switch(t)
  {
    case PlyType::Char:
      return static_cast<float>(loadScalar<int8_t>(p, swap));
    case PlyType::UChar:
      return static_cast<float>(loadScalar<uint8_t>(p, swap));
    case PlyType::Short:
      return static_cast<float>(loadScalar<int16_t>(p, swap));
    case PlyType::UShort:
      return static_cast<float>(loadScalar<uint16_t>(p, swap));
    case PlyType::Int:
      return static_cast<float>(loadScalar<int32_t>(p, swap));
    case PlyType::UInt:
      return static_cast<float>(loadScalar<uint32_t>(p, swap));
    case PlyType::Float:
      return loadScalar<float>(p, swap);
    case PlyType::Double:
      return static_cast<float>(loadScalar<double>(p, swap));
    default:
      return 0.0f;
  }
}

uint64_t scalarToCount(const uint8_t* p, PlyType t, bool swap)
{
  switch(t)
  {
    case PlyType::Char:
      return static_cast<uint64_t>(loadScalar<int8_t>(p, swap));
    case PlyType::UChar:
      return loadScalar<uint8_t>(p, swap);
    case PlyType::Short:
      return static_cast<uint64_t>(loadScalar<int16_t>(p, swap));
    case PlyType::UShort:
      return loadScalar<uint16_t>(p, swap);
    case PlyType::Int:
      return static_cast<uint64_t>(loadScalar<int32_t>(p, swap));
    case PlyType::UInt:
      return loadScalar<uint32_t>(p, swap);
    default:
      return 0;
  }
}

bool loadPly(const std::string& path, vkgs_scene& out)
{
  FileBytes data;
  if(!data.open(path))
    return false;
  // ---- header ----
  size_t pos = 0;
  auto   nextLine = [&](std::string& line) -> bool {
    if(pos >= data.size())
      return false;
    size_t e = pos;
    while(e < data.size() && data[e] != '\n')
      e++;
    line.assign(reinterpret_cast<const char*>(data.data()) + pos, e - pos);
    if(!line.empty() && line.back() == '\r')
      line.pop_back();
    pos = e + 1;
    return true;
  };
  std::string line;
  if(!nextLine(line) || line != "ply")
    return failLoad("not a ply file: " + path);
  enum
  {
    Ascii,
    Little,
    Big
  } format       = Ascii;
  bool haveFormat = false;
  std::vector<PlyElement> elements;
  while(true)
  {
    if(!nextLine(line))
      return failLoad("ply header not terminated");
    std::istringstream ss(line);
    std::string        kw;
    ss >> kw;
    if(kw == "end_header")
      break;
    if(kw == "format")
    {
      std::string f;
      ss >> f;
      if(f == "ascii")
        format = Ascii;
      else if(f == "binary_little_endian")
        format = Little;
      else if(f == "binary_big_endian")
        format = Big;
      else
        return failLoad("unknown ply format " + f);
      haveFormat = true;
    }
    else if(kw == "element")
    {
      PlyElement e;
      ss >> e.name >> e.count;
      elements.push_back(e);
    }
    else if(kw == "property")
    {
      if(elements.empty())
        return failLoad("ply property before any element");
      PlyProp     p;
      std::string t;
      ss >> t;
      if(t == "list")
      {
        std::string ct, vt;
        ss >> ct >> vt >> p.name;
        p.countType = parseType(ct);
        p.type      = parseType(vt);
        if(p.countType == PlyType::None)
          return failLoad("bad ply list count type");
        elements.back().fixedSize = false;
      }
      else
      {
        p.type = parseType(t);
        ss >> p.name;
      }
      if(p.type == PlyType::None)
        return failLoad("unknown ply property type in: " + line);
      p.offset = elements.back().rowBytes;
      elements.back().rowBytes += typeSize(p.type);
      elements.back().props.push_back(p);
    }
    // comment / obj_info: ignored
  }
  if(!haveFormat)
    return failLoad("ply header has no format line");
  const bool swap = format == Big;  // host is little endian

  // ---- find the first non-empty vertex element; skip what precedes it ----
  for(const PlyElement& el : elements)
  {
    const bool isVertex = el.name == "vertex";
    if(isVertex && el.count > 0)
    {
      if(!el.fixedSize)
        return failLoad("vertex element with list properties is not a 3DGS ply");
      const size_t n = el.count;
      // the rows as floats, by property index
      const size_t         np = el.props.size();
      // Validate the count BEFORE any multiplication (a crafted header could wrap n * rowBytes or n * np modulo 2^64):
      // the upload refuses more than 2^31-1 splats anyway, a binary file must hold n whole rows, and an ascii file
      // needs at least two bytes per value.
      if(n > 0x7fffffffull || np == 0)
        return failLoad("ply vertex count out of range");
      if(format != Ascii && (el.rowBytes == 0 || pos > data.size() || n > (data.size() - pos) / el.rowBytes))
        return failLoad("binary ply ends early");
      if(format == Ascii && (pos > data.size() || n > (data.size() - pos) / (2 * np) + 1))
        return failLoad("ascii ply ends early");
      std::vector<float>   rows;
      const uint8_t*       base = nullptr;
      if(format == Ascii)
      {
        rows.resize(n * np);
        const char* p   = reinterpret_cast<const char*>(data.data()) + pos;
        const char* end = reinterpret_cast<const char*>(data.data()) + data.size();
        std::string tok;
        for(size_t i = 0; i < n * np; i++)
        {
          while(p < end && std::isspace(static_cast<unsigned char>(*p)))
            p++;
          const char* q = p;
          while(q < end && !std::isspace(static_cast<unsigned char>(*q)))
            q++;
          if(p == q)
            return failLoad("ascii ply ends early");
          tok.assign(p, q - p);
          const PlyType t = el.props[i % np].type;
          // miniply parses ascii values with the property's own type, then converts to float
          if(t == PlyType::Float)
            rows[i] = std::strtof(tok.c_str(), nullptr);
          else if(t == PlyType::Double)
            rows[i] = static_cast<float>(std::strtod(tok.c_str(), nullptr));
          else
            rows[i] = static_cast<float>(std::strtoll(tok.c_str(), nullptr, 10));
          p = q;
        }
      }
      else
      {
        base = data.data() + pos;  // (n rows fit: checked above)
      }
      auto find = [&](const char* name) -> int {
        for(size_t k = 0; k < np; k++)
          if(el.props[k].name == name)
            return static_cast<int>(k);
        return -1;
      };
      // extract a group of properties, all-or-nothing like miniply::find_properties
      auto extract = [&](const std::vector<std::string>& names, std::vector<float>& dst) -> bool {
        std::vector<int> idx;
        for(const std::string& nm : names)
        {
          const int k = find(nm.c_str());
          if(k < 0)
            return false;
          idx.push_back(k);
        }
        dst.resize(n * names.size());
        const size_t m = idx.size();
        bool         plainFloats = format != Ascii && !swap;  // the INRIA layout: native-endian float32 properties
        for(size_t j = 0; j < m; j++)
          plainFloats = plainFloats && el.props[idx[j]].type == PlyType::Float;
        if(plainFloats)
        {
          std::vector<size_t> off(m);
          for(size_t j = 0; j < m; j++)
            off[j] = el.props[idx[j]].offset;
          const size_t rowBytes = el.rowBytes;
          float*       d        = dst.data();
          parallelFor(n, [&](size_t b, size_t e) {
            for(size_t i = b; i < e; i++)
            {
              const uint8_t* row = base + i * rowBytes;
              for(size_t j = 0; j < m; j++)
                std::memcpy(d + i * m + j, row + off[j], 4);
            }
          });
          return true;
        }
        for(size_t i = 0; i < n; i++)
          for(size_t j = 0; j < m; j++)
          {
            const PlyProp& pr = el.props[idx[j]];
            dst[i * m + j] = format == Ascii ? rows[i * np + idx[j]] : scalarToFloat(base + i * el.rowBytes + pr.offset, pr.type, swap);
          }
        return true;
      };
      std::vector<std::string> rest;
      for(int k = 0; k < 45; k++)
        rest.push_back("f_rest_" + std::to_string(k));
      extract(rest, out.f_rest);
      const bool ok = extract({"x", "y", "z"}, out.positions) & extract({"opacity"}, out.opacity)
                      & extract({"scale_0", "scale_1", "scale_2"}, out.scale) & extract({"rot_0", "rot_1", "rot_2", "rot_3"}, out.rotation)
                      & extract({"f_dc_0", "f_dc_1", "f_dc_2"}, out.f_dc);
      if(!ok)
        return failLoad("ply vertex element lacks 3DGS properties (x y z opacity scale_* rot_* f_dc_*)");
      out.fRestPerSplat = out.f_rest.empty() ? 0u : 45u;
      rdfToRub(out);
      return true;
    }
    // skip this element
    if(format == Ascii)
    {
      std::string skip;
      for(size_t i = 0; i < el.count; i++)
        if(!nextLine(skip))
          return failLoad("ascii ply ends early");
    }
    else if(el.fixedSize)
    {
      // checked skip: count * rowBytes must stay inside the file (and must not wrap)
      if(pos > data.size() || (el.rowBytes != 0 && el.count > (data.size() - pos) / el.rowBytes))
        return failLoad("binary ply ends early");
      pos += el.count * el.rowBytes;
    }
    else
    {
      for(size_t i = 0; i < el.count; i++)
      {
        for(const PlyProp& pr : el.props)
        {
          if(pr.countType == PlyType::None)
            pos += typeSize(pr.type);
          else
          {
            if(pos > data.size() || typeSize(pr.countType) > data.size() - pos)
              return failLoad("binary ply ends early");
            const uint64_t c = scalarToCount(data.data() + pos, pr.countType, swap);
            if(c > (data.size() - pos) / typeSize(pr.type))
              return failLoad("binary ply ends early");
            pos += typeSize(pr.countType) + c * typeSize(pr.type);
          }
        }
        if(pos > data.size())
          return failLoad("binary ply ends early");
      }
    }
    if(pos > data.size())
      return failLoad("binary ply ends early");
  }
  return failLoad("invalid 3DGS PLY file (no vertex element)");
}

// ---------------------------------------------------------------------------------------------
// .splat

bool loadSplat(const std::string& path, vkgs_scene& out)
{
  std::vector<uint8_t> data;
  if(!readFile(path, data))
    return false;
  if(data.size() % 32 != 0)
    return failLoad("invalid .splat file size (not a multiple of 32 bytes)");
  const size_t n = data.size() / 32;
  if(n == 0)
    return failLoad("empty .splat file");
  out.positions.resize(n * 3);
  out.scale.resize(n * 3);
  out.rotation.resize(n * 4);
  out.opacity.resize(n);
  out.f_dc.resize(n * 3);
  out.f_rest.clear();
  out.fRestPerSplat     = 0;
  constexpr float SH_C0 = 0.28209479177387814f;
  for(size_t i = 0; i < n; i++)
  {
    const uint8_t* r = data.data() + 32 * i;
    float          f[6];
    std::memcpy(f, r, 24);
    const uint8_t* color = r + 24;
    const uint8_t* rot   = r + 28;
    for(int k = 0; k < 3; k++)
    {
      out.positions[3 * i + k] = f[k];
      out.scale[3 * i + k]     = std::log(f[3 + k]);
      out.f_dc[3 * i + k]      = (color[k] / 255.0f - 0.5f) / SH_C0;
    }
    for(int k = 0; k < 4; k++)
      out.rotation[4 * i + k] = (static_cast<float>(rot[k]) - 128.0f) / 128.0f;
    const float alpha        = color[3] / 255.0f;
    const float alphaClamped = std::clamp(alpha, 1e-6f, 1.0f - 1e-6f);
    out.opacity[i]           = -std::log((1.0f / alphaClamped) - 1.0f);
  }
  rdfToRub(out);
  return true;
}

// ---------------------------------------------------------------------------------------------
// .spz

bool gunzip(const std::vector<uint8_t>& in, std::vector<uint8_t>& out)
{
  z_stream zs{};
  zs.next_in  = const_cast<Bytef*>(in.data());
  zs.avail_in = static_cast<uInt>(in.size());
  if(inflateInit2(&zs, 16 | MAX_WBITS) != Z_OK)
    return false;
  std::vector<uint8_t> buf(1 << 16);
  bool                 ok = false;
  out.clear();
  while(true)
  {
    zs.next_out  = buf.data();
    zs.avail_out = static_cast<uInt>(buf.size());
    const int rc = inflate(&zs, Z_NO_FLUSH);
    if(rc != Z_OK && rc != Z_STREAM_END)
      break;
    out.insert(out.end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
    if(rc == Z_STREAM_END)
    {
      ok = true;
      break;
    }
  }
  inflateEnd(&zs);
  return ok;
}

float halfBitsToFloat(uint16_t h)
{
  // spz halfToFloat (load-spz.cc): sign/exponent/mantissa expansion incl. subnormals
  const uint32_t sign = (h >> 15) & 1u, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
  const float    sgn  = sign ? -1.0f : 1.0f;
  if(exp == 0)
    return sgn * std::pow(2.0f, -14.0f) * static_cast<float>(man) / 1024.0f;
  return sgn * std::pow(2.0f, static_cast<float>(exp) - 15.0f) * (1.0f + static_cast<float>(man) / 1024.0f);
}

bool loadSpz(const std::string& path, vkgs_scene& out)
{
  std::vector<uint8_t> gz, raw;
  if(!readFile(path, gz))
    return false;
  if(!gunzip(gz, raw))
    return failLoad("spz: gzip decode failed");
  if(raw.size() < 16)
    return failLoad("spz: header not found");
  uint32_t magic, version, numPoints;
  std::memcpy(&magic, raw.data(), 4);
  std::memcpy(&version, raw.data() + 4, 4);
  std::memcpy(&numPoints, raw.data() + 8, 4);
  const uint8_t shDegree = raw[12], fractionalBits = raw[13];
  if(magic != 0x5053474eu)
    return failLoad("spz: header not found");
  if(version < 1 || version > 3)
    return failLoad("spz: version not supported");
  if(numPoints > 10000000u)
    return failLoad("spz: too many points");
  if(shDegree > 3)
    return failLoad("spz: unsupported SH degree");
  static const int dims[4] = {0, 3, 8, 15};
  const size_t     n = numPoints, shDim = static_cast<size_t>(dims[shDegree]);
  const bool       f16 = version == 1, smallest3 = version >= 3;
  const size_t     posBytes = n * 3 * (f16 ? 2 : 3), rotBytes = n * (smallest3 ? 4 : 3);
  const size_t     need     = 16 + posBytes + n + n * 3 + n * 3 + rotBytes + n * shDim * 3;
  if(raw.size() < need)
    return failLoad("spz: read error");
  if(n == 0)
    return failLoad("spz: empty cloud");
  const uint8_t* pPos = raw.data() + 16;
  const uint8_t* pAlp = pPos + posBytes;
  const uint8_t* pCol = pAlp + n;
  const uint8_t* pScl = pCol + n * 3;
  const uint8_t* pRot = pScl + n * 3;
  const uint8_t* pSh  = pRot + rotBytes;

  out.positions.resize(n * 3);
  out.scale.resize(n * 3);
  out.rotation.resize(n * 4);
  out.opacity.resize(n);
  out.f_dc.resize(n * 3);
  out.f_rest.resize(n * shDim * 3);
  out.fRestPerSplat = static_cast<uint32_t>(shDim * 3);

  if(f16)
  {
    parallelFor(n * 3, [&](size_t b, size_t e) {
      for(size_t i = b; i < e; i++)
      {
        uint16_t h;
        std::memcpy(&h, pPos + 2 * i, 2);
        out.positions[i] = halfBitsToFloat(h);
      }
    });
  }
  else
  {
    const float scale = static_cast<float>(1.0 / (1 << fractionalBits));
    parallelFor(n * 3, [&](size_t b, size_t e) {
      for(size_t i = b; i < e; i++)
      {
        int32_t fixed32 = pPos[i * 3 + 0];
        fixed32 |= pPos[i * 3 + 1] << 8;
        fixed32 |= pPos[i * 3 + 2] << 16;
        fixed32 |= (fixed32 & 0x800000) ? 0xff000000 : 0;
        out.positions[i] = static_cast<float>(fixed32) * scale;
      }
    });
  }
  parallelFor(n * 3, [&](size_t b, size_t e) {
    for(size_t i = b; i < e; i++)
      out.scale[i] = pScl[i] / 16.0f - 10.0f;
  });
  constexpr float sqrt1_2 = static_cast<float>(0.707106781186547524401);
  parallelFor(n, [&](size_t rb, size_t re) {
  for(size_t i = rb; i < re; i++)
  {
    float q[4];  // x y z w as spz stores them
    if(smallest3)
    {
      const uint8_t* r = pRot + 4 * i;
      uint32_t comp = static_cast<uint32_t>(r[0]) + (static_cast<uint32_t>(r[1]) << 8) + (static_cast<uint32_t>(r[2]) << 16)
                      + (static_cast<uint32_t>(r[3]) << 24);
      constexpr uint32_t cMask    = (1u << 9u) - 1u;
      const int          iLargest = static_cast<int>(comp >> 30);
      float              sumSquares = 0;
      for(int k = 3; k >= 0; --k)
      {
        if(k != iLargest)
        {
          const uint32_t mag    = comp & cMask;
          const uint32_t negbit = (comp >> 9u) & 0x1u;
          comp                  = comp >> 10u;
          q[k]                  = sqrt1_2 * static_cast<float>(mag) / static_cast<float>(cMask);
          if(negbit == 1)
            q[k] = -q[k];
          sumSquares += q[k] * q[k];
        }
      }
      q[iLargest] = std::sqrt(1.0f - sumSquares);
    }
    else
    {
      const uint8_t* r = pRot + 3 * i;
      for(int k = 0; k < 3; k++)
        q[k] = (static_cast<float>(r[k]) * (1.0f / 127.5f) + -1.0f) * 1.0f;  // flipQ = 1 (RUB -> RUB)
      q[3] = std::sqrt(std::max(0.0f, 1.0f - (q[0] * q[0] + q[1] * q[1] + q[2] * q[2])));
    }
    // ply_loader_async.cpp:314-321: xyzw -> wxyz
    out.rotation[4 * i + 0] = q[3];
    out.rotation[4 * i + 1] = q[0];
    out.rotation[4 * i + 2] = q[1];
    out.rotation[4 * i + 3] = q[2];
  }
  });
  constexpr float colorScale = 0.15f;
  parallelFor(n, [&](size_t b, size_t e) {
    for(size_t i = b; i < e; i++)
    {
      const float a  = pAlp[i] / 255.0f;
      out.opacity[i] = std::log(a / (1.0f - a));  // invSigmoid
      for(size_t c = 0; c < 3; c++)
        out.f_dc[3 * i + c] = ((pCol[3 * i + c] / 255.0f) - 0.5f) / colorScale;
      // SH: spz keeps RGB inner per coefficient; SplatSet wants channel-major per splat (:326-346)
      for(size_t j = 0; j < shDim; j++)
        for(size_t c = 0; c < 3; c++)
          out.f_rest[i * shDim * 3 + c * shDim + j] = (static_cast<float>(pSh[(i * shDim + j) * 3 + c]) - 128.0f) / 128.0f;
    }
  });
  return true;
}

}  // namespace

extern "C" {

int vkgs_scene_load(const char* path, vkgs_scene** out)
{
  if(!path || !out)
    return VKGS_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  g_loaderError.clear();
  vkgs_scene*       s   = new vkgs_scene();
  const std::string p   = path;
  const std::string ext = lowerExt(p);
  bool              ok  = false;
  try  // no exception crosses the C ABI (a header that promises 2^40 vertices ends in bad_alloc, not in terminate)
  {
    if(ext == ".splat")
      ok = loadSplat(p, *s);
    else if(ext == ".spz")
      ok = loadSpz(p, *s);
    else
      ok = loadPly(p, *s);
  }
  catch(const std::exception& e)
  {
    ok = failLoad(std::string("loading ") + p + " failed: " + e.what());
  }
  if(!ok)
  {
    delete s;
    return VKGS_ERR_IO;
  }
  *out = s;
  return VKGS_OK;
}

const char* vkgs_scene_load_error(void)
{
  return g_loaderError.c_str();
}

int vkgs_scene_view(const vkgs_scene* s, vkgs_splat_set_view* view)
{
  if(!s || !view)
    return VKGS_ERR_INVALID_ARGUMENT;
  view->positions        = s->positions.data();
  view->f_dc             = s->f_dc.data();
  view->f_rest           = s->f_rest.empty() ? nullptr : s->f_rest.data();
  view->opacity          = s->opacity.data();
  view->scale            = s->scale.data();
  view->rotation         = s->rotation.data();
  view->count            = s->positions.size() / 3;
  view->f_rest_per_splat = s->fRestPerSplat;
  view->_pad             = 0;
  return VKGS_OK;
}

int vkgs_scene_free(vkgs_scene* s)
{
  delete s;
  return VKGS_OK;
}

}  // extern "C"
