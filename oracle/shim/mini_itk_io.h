/*
 * mini_itk_io.h -- stand-ins for the ITK I/O and utility filters that the reference's TEST PROGRAMS use
 * (/root/reference/test/itk2DDiffusionTest_{GS,WJ}.cxx, itkVEDTest_GS.cxx): ImageFileReader, ImageFileWriter, CastImageFilter,
 * ChangeInformationImageFilter.  TEST INFRASTRUCTURE ONLY: with them those programs compile UNMODIFIED against this repo's
 * drop-in headers (include/itk*.h) and run on the B200 path (tests/test_ref_tests_dropin.py).  Not ITK code.
 *
 * File formats: MetaImage (.mhd + .raw / zlib .zraw) is read and written directly.  Any other file name (the 2-D tests use
 * test_data/lena.jpg) is served through a MetaImage side-car "<name>.mhd" that the test harness prepares / collects -- this
 * image has no JPEG library, and the decoder is not part of what is being tested.
 */
#ifndef MINI_ITK_IO_H
#define MINI_ITK_IO_H

#include <zlib.h>

#include <cstdint>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#include "mini_itk.h"

namespace itk
{
namespace shim_io
{
inline bool ends_with(const std::string& s, const std::string& e) { return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0; }
inline std::string header_path(const std::string& name) { return ends_with(name, ".mhd") ? name : name + ".mhd"; }
inline std::string dir_of(const std::string& p) { const size_t i = p.find_last_of('/'); return i == std::string::npos ? std::string() : p.substr(0, i + 1); }
inline std::string base_of(const std::string& p) { const size_t i = p.find_last_of('/'); return i == std::string::npos ? p : p.substr(i + 1); }

template <typename TSrc, typename TImage>
void convert(const std::vector<char>& raw, TImage* img)
{
  const size_t n = img->GetLargestPossibleRegion().GetNumberOfPixels();
  if (raw.size() < n * sizeof(TSrc)) throw std::runtime_error("ImageFileReader stand-in: data file too short");
  const TSrc* s = reinterpret_cast<const TSrc*>(raw.data());
  for (size_t i = 0; i < n; ++i) img->GetBufferPointer()[i] = static_cast<typename TImage::PixelType>(s[i]);
}
template <typename T> struct MetName;
template <> struct MetName<unsigned char> { static const char* value() { return "MET_UCHAR"; } };
template <> struct MetName<short> { static const char* value() { return "MET_SHORT"; } };
template <> struct MetName<float> { static const char* value() { return "MET_FLOAT"; } };
template <> struct MetName<double> { static const char* value() { return "MET_DOUBLE"; } };
}  // namespace shim_io

template <typename TImage>
class ImageFileReader : public LightObject
{
public:
  typedef ImageFileReader Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro(Self);
  void SetFileName(const std::string& n) { m_Name = n; }
  TImage* GetOutput() { return m_Output.GetPointer(); }
  void Update()
  {
    const unsigned int D = TImage::ImageDimension;
    const std::string hp = shim_io::header_path(m_Name);
    std::ifstream h(hp.c_str());
    if (!h) throw std::runtime_error("ImageFileReader stand-in: cannot open " + hp);
    std::map<std::string, std::string> kv;
    std::string line;
    while (std::getline(h, line)) {
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      auto trim = [](std::string s) { const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r"); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
      kv[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
    if (static_cast<unsigned int>(std::atoi(kv["NDims"].c_str())) != D) throw std::runtime_error("ImageFileReader stand-in: NDims mismatch in " + hp);
    typename TImage::IndexType idx;
    typename TImage::SizeType size;
    typename TImage::SpacingType sp;
    typename TImage::PointType org;
    typename TImage::DirectionType dir;
    idx.Fill(0);
    sp.Fill(1.0);
    org.Fill(0.0);
    {
      std::istringstream a(kv["DimSize"]);
      for (unsigned int d = 0; d < D; ++d) { unsigned long v = 0; a >> v; size[d] = v; }
      if (kv.count("ElementSpacing")) { std::istringstream b(kv["ElementSpacing"]); for (unsigned int d = 0; d < D; ++d) b >> sp[d]; }
      if (kv.count("Offset")) { std::istringstream b(kv["Offset"]); for (unsigned int d = 0; d < D; ++d) b >> org[d]; }
    }
    m_Output = TImage::New();
    dir = m_Output->GetDirection();
    if (kv.count("TransformMatrix")) { std::istringstream b(kv["TransformMatrix"]); for (unsigned int d = 0; d < D * D; ++d) b >> dir[d]; }
    m_Output->SetRegions(typename TImage::RegionType(idx, size));
    m_Output->Allocate();
    m_Output->SetSpacing(sp);
    m_Output->SetOrigin(org);
    m_Output->SetDirection(dir);
    const std::string dp = shim_io::dir_of(hp) + kv["ElementDataFile"];
    std::ifstream f(dp.c_str(), std::ios::binary);
    if (!f) throw std::runtime_error("ImageFileReader stand-in: cannot open " + dp);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const std::string et = kv["ElementType"];
    const size_t es = et == "MET_UCHAR" ? 1 : et == "MET_SHORT" ? 2 : et == "MET_FLOAT" ? 4 : et == "MET_DOUBLE" ? 8 : 0;
    if (!es) throw std::runtime_error("ImageFileReader stand-in: unsupported ElementType " + et);
    if (kv["CompressedData"] == "True") {
      std::vector<char> out(m_Output->GetLargestPossibleRegion().GetNumberOfPixels() * es);
      uLongf len = static_cast<uLongf>(out.size());
      if (uncompress(reinterpret_cast<Bytef*>(out.data()), &len, reinterpret_cast<const Bytef*>(raw.data()), static_cast<uLong>(raw.size())) != Z_OK)
        throw std::runtime_error("ImageFileReader stand-in: zlib failure on " + dp);
      raw.swap(out);
    }
    if (es == 1) shim_io::convert<unsigned char>(raw, m_Output.GetPointer());
    else if (es == 2) shim_io::convert<short>(raw, m_Output.GetPointer());
    else if (es == 4) shim_io::convert<float>(raw, m_Output.GetPointer());
    else shim_io::convert<double>(raw, m_Output.GetPointer());
  }
protected:
  ImageFileReader() {}
private:
  std::string m_Name;
  typename TImage::Pointer m_Output;
};

template <typename TImage>
class ImageFileWriter : public LightObject
{
public:
  typedef ImageFileWriter Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro(Self);
  void SetFileName(const std::string& n) { m_Name = n; }
  void SetInput(const TImage* img) { m_Input = img; }
  void Update()
  {
    const unsigned int D = TImage::ImageDimension;
    if (!m_Input) throw std::runtime_error("ImageFileWriter stand-in: no input");
    const std::string hp = shim_io::header_path(m_Name);
    std::string stem = shim_io::base_of(hp);
    stem = stem.substr(0, stem.size() - 4);
    const std::string data = stem + ".raw";
    const typename TImage::RegionType r = m_Input->GetLargestPossibleRegion();
    std::ofstream h(hp.c_str());
    if (!h) throw std::runtime_error("ImageFileWriter stand-in: cannot create " + hp);
    h.precision(17);
    h << "ObjectType = Image\nNDims = " << D << "\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = False\nTransformMatrix =";
    for (unsigned int d = 0; d < D * D; ++d) h << " " << m_Input->GetDirection()[d];
    h << "\nOffset =";
    for (unsigned int d = 0; d < D; ++d) h << " " << m_Input->GetOrigin()[d];
    h << "\nElementSpacing =";
    for (unsigned int d = 0; d < D; ++d) h << " " << m_Input->GetSpacing()[d];
    h << "\nDimSize =";
    for (unsigned int d = 0; d < D; ++d) h << " " << r.GetSize(d);
    h << "\nElementType = " << shim_io::MetName<typename TImage::PixelType>::value() << "\nElementDataFile = " << data << "\n";
    std::ofstream f((shim_io::dir_of(hp) + data).c_str(), std::ios::binary);
    f.write(reinterpret_cast<const char*>(m_Input->GetBufferPointer()), static_cast<std::streamsize>(r.GetNumberOfPixels() * sizeof(typename TImage::PixelType)));
    if (!f) throw std::runtime_error("ImageFileWriter stand-in: cannot write " + data);
  }
protected:
  ImageFileWriter() : m_Input(nullptr) {}
private:
  std::string m_Name;
  const TImage* m_Input;
};

template <typename TIn, typename TOut>
class CastImageFilter : public LightObject
{
public:
  typedef CastImageFilter Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro(Self);
  void SetInput(const TIn* in) { m_Input = in; }
  TOut* GetOutput() { return m_Output.GetPointer(); }
  void Update()
  {
    m_Output = TOut::New();
    m_Output->SetRegions(m_Input->GetLargestPossibleRegion());
    m_Output->Allocate();
    m_Output->SetSpacing(m_Input->GetSpacing());
    m_Output->SetOrigin(m_Input->GetOrigin());
    m_Output->SetDirection(m_Input->GetDirection());
    const size_t n = m_Input->GetLargestPossibleRegion().GetNumberOfPixels();
    for (size_t i = 0; i < n; ++i) m_Output->GetBufferPointer()[i] = static_cast<typename TOut::PixelType>(m_Input->GetBufferPointer()[i]);
  }
protected:
  CastImageFilter() : m_Input(nullptr) {}
private:
  const TIn* m_Input;
  typename TOut::Pointer m_Output;
};

template <typename TImage>
class ChangeInformationImageFilter : public LightObject
{
public:
  typedef ChangeInformationImageFilter Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro(Self);
  void SetInput(TImage* in) { m_Image = in; }
  void SetOutputDirection(const typename TImage::DirectionType& d) { m_Direction = d; }
  void ChangeDirectionOn() { m_Change = true; }
  void UpdateOutputInformation() { if (m_Change && m_Image) m_Image->SetDirection(m_Direction); }
  void Update() { UpdateOutputInformation(); }
  TImage* GetOutput() { return m_Image; }
protected:
  ChangeInformationImageFilter() : m_Image(nullptr), m_Change(false) {}
private:
  TImage* m_Image;
  typename TImage::DirectionType m_Direction;
  bool m_Change;
};
}  // namespace itk

#endif  // MINI_ITK_IO_H
