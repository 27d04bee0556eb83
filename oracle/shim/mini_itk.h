/*
 * mini_itk.h -- a minimal stand-in for the parts of ITK and VXL/vnl that the reference's solver headers
 * (/root/reference/include/itkMultigridAnisotropicDiffusionImageFilter.{h,hxx} and the include/mad headers) use, so
 * that those headers compile UNMODIFIED, where they lie, into oracle/_ref/libmadref.so (oracle/Makefile,
 * target `ref`).  TEST INFRASTRUCTURE ONLY: it exists to pin oracle/mad_oracle.c against the reference's own
 * code; nothing in the product links it.
 *
 * Written from the ITK public API as the reference uses it (SURVEY.md section 8c lists the surface); it is
 * not ITK code.  Semantics that matter for the numerics and are reproduced on purpose:
 *   - raster (x fastest) iteration order of region and neighbourhood iterators,
 *   - Neighborhood storage order and GetOffset(i),
 *   - neighbourhood iterators reading the live image (Gauss-Seidel sees its own updates),
 *   - ImageBoundaryFacesCalculator: interior region first, then non-overlapping boundary faces,
 *   - SymmetricSecondRankTensor upper-triangular row-major storage.
 * vnl_sparse_lu is replaced by a dense LU with partial pivoting (an exact solve either way).
 */
#ifndef MINI_ITK_H
#define MINI_ITK_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace itk
{
typedef long IndexValueType;
typedef unsigned long SizeValueType;
typedef long OffsetValueType;

template <unsigned int D> struct Size;

template <unsigned int D>
struct Offset {
  OffsetValueType m_v[D];
  OffsetValueType& operator[](unsigned int d) { return m_v[d]; }
  OffsetValueType operator[](unsigned int d) const { return m_v[d]; }
  void Fill(OffsetValueType v) { for (unsigned int d = 0; d < D; ++d) m_v[d] = v; }
  Offset operator+(const Offset& o) const { Offset r; for (unsigned int d = 0; d < D; ++d) r.m_v[d] = m_v[d] + o.m_v[d]; return r; }
  Offset operator-(const Offset& o) const { Offset r; for (unsigned int d = 0; d < D; ++d) r.m_v[d] = m_v[d] - o.m_v[d]; return r; }
  Offset operator+(const Size<D>& s) const;
  Offset& operator+=(const Offset& o) { for (unsigned int d = 0; d < D; ++d) m_v[d] += o.m_v[d]; return *this; }
  Offset& operator-=(const Offset& o) { for (unsigned int d = 0; d < D; ++d) m_v[d] -= o.m_v[d]; return *this; }
  bool operator==(const Offset& o) const { for (unsigned int d = 0; d < D; ++d) if (m_v[d] != o.m_v[d]) return false; return true; }
  bool operator!=(const Offset& o) const { return !(*this == o); }
};

template <unsigned int D>
struct Size {
  SizeValueType m_v[D];
  SizeValueType& operator[](unsigned int d) { return m_v[d]; }
  SizeValueType operator[](unsigned int d) const { return m_v[d]; }
  void Fill(SizeValueType v) { for (unsigned int d = 0; d < D; ++d) m_v[d] = v; }
  bool operator==(const Size& o) const { for (unsigned int d = 0; d < D; ++d) if (m_v[d] != o.m_v[d]) return false; return true; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};

template <unsigned int D>
Offset<D> Offset<D>::operator+(const Size<D>& s) const
{
  Offset r;
  for (unsigned int d = 0; d < D; ++d) r.m_v[d] = m_v[d] + static_cast<OffsetValueType>(s[d]);
  return r;
}

template <unsigned int D>
struct Index {
  IndexValueType m_v[D];
  IndexValueType& operator[](unsigned int d) { return m_v[d]; }
  IndexValueType operator[](unsigned int d) const { return m_v[d]; }
  void Fill(IndexValueType v) { for (unsigned int d = 0; d < D; ++d) m_v[d] = v; }
  Index operator+(const Offset<D>& o) const { Index r; for (unsigned int d = 0; d < D; ++d) r.m_v[d] = m_v[d] + o[d]; return r; }
  Index operator-(const Offset<D>& o) const { Index r; for (unsigned int d = 0; d < D; ++d) r.m_v[d] = m_v[d] - o[d]; return r; }
  bool operator==(const Index& o) const { for (unsigned int d = 0; d < D; ++d) if (m_v[d] != o.m_v[d]) return false; return true; }
  bool operator!=(const Index& o) const { return !(*this == o); }
};

template <typename T, unsigned int D>
struct FixedVector {
  T m_v[D];
  FixedVector() { for (unsigned int d = 0; d < D; ++d) m_v[d] = T(); }
  T& operator[](unsigned int d) { return m_v[d]; }
  const T& operator[](unsigned int d) const { return m_v[d]; }
  void Fill(T v) { for (unsigned int d = 0; d < D; ++d) m_v[d] = v; }
};

template <unsigned int D>
class ImageRegion
{
public:
  typedef Index<D> IndexType;
  typedef Size<D> SizeType;
  ImageRegion() { m_Index.Fill(0); m_Size.Fill(0); }
  ImageRegion(const IndexType& i, const SizeType& s) : m_Index(i), m_Size(s) {}
  const SizeType& GetSize() const { return m_Size; }
  SizeValueType GetSize(unsigned int d) const { return m_Size[d]; }
  const IndexType& GetIndex() const { return m_Index; }
  IndexValueType GetIndex(unsigned int d) const { return m_Index[d]; }
  void SetSize(const SizeType& s) { m_Size = s; }
  void SetIndex(const IndexType& i) { m_Index = i; }
  SizeValueType GetNumberOfPixels() const { SizeValueType n = 1; for (unsigned int d = 0; d < D; ++d) n *= m_Size[d]; return n; }
  bool IsInside(const IndexType& i) const
  {
    for (unsigned int d = 0; d < D; ++d)
      if (i[d] < m_Index[d] || i[d] >= m_Index[d] + static_cast<IndexValueType>(m_Size[d])) return false;
    return true;
  }
private:
  IndexType m_Index;
  SizeType m_Size;
};

// ---- reference counting ------------------------------------------------------------------------
class LightObject
{
public:
  LightObject() : m_RefCount(0) {}
  virtual ~LightObject() {}
  void Register() const { ++m_RefCount; }
  void UnRegister() const { if (--m_RefCount <= 0) delete this; }
  virtual const char* GetNameOfClass() const { return "LightObject"; }
  virtual void Modified() const { ++m_MTime; }  // itk::Object's modification stamp (what itkSetMacro bumps in real ITK)
  unsigned long GetMTime() const { return m_MTime; }
private:
  mutable long m_RefCount;
  mutable unsigned long m_MTime = 0;
};

template <typename T>
class SmartPointer
{
public:
  SmartPointer() : m_P(nullptr) {}
  SmartPointer(T* p) : m_P(p) { if (m_P) m_P->Register(); }
  SmartPointer(const SmartPointer& o) : m_P(o.m_P) { if (m_P) m_P->Register(); }
  template <typename U> SmartPointer(const SmartPointer<U>& o) : m_P(o.GetPointer()) { if (m_P) m_P->Register(); }
  ~SmartPointer() { if (m_P) m_P->UnRegister(); }
  SmartPointer& operator=(const SmartPointer& o) { return *this = o.m_P; }
  SmartPointer& operator=(T* p)
  {
    if (p) p->Register();
    if (m_P) m_P->UnRegister();
    m_P = p;
    return *this;
  }
  T* operator->() const { return m_P; }
  T& operator*() const { return *m_P; }
  operator T*() const { return m_P; }
  T* GetPointer() const { return m_P; }
  bool IsNull() const { return m_P == nullptr; }
  bool IsNotNull() const { return m_P != nullptr; }
private:
  T* m_P;
};

#define itkNewMacro(x) \
  static Pointer New() { Pointer p = new x; return p; }
#define itkTypeMacro(thisClass, superclass) \
  virtual const char* GetNameOfClass() const { return #thisClass; }
#define itkSetMacro(name, type) \
  virtual void Set##name(const type _arg) { this->m_##name = _arg; }
#define itkGetConstMacro(name, type) \
  virtual type Get##name() const { return this->m_##name; }
#define itkGetMacro(name, type) \
  virtual type Get##name() { return this->m_##name; }

// ---- pixel containers --------------------------------------------------------------------------
template <typename T, unsigned int D>
class SymmetricSecondRankTensor
{
public:
  enum { InternalDimension = D * (D + 1) / 2 };
  SymmetricSecondRankTensor() { Fill(T()); }
  void Fill(const T& v) { for (unsigned int k = 0; k < InternalDimension; ++k) m_v[k] = v; }
  T& operator()(unsigned int r, unsigned int c) { return m_v[Pos(r, c)]; }
  const T& operator()(unsigned int r, unsigned int c) const { return m_v[Pos(r, c)]; }
  T& operator[](unsigned int k) { return m_v[k]; }
  const T& operator[](unsigned int k) const { return m_v[k]; }
private:
  static unsigned int Pos(unsigned int r, unsigned int c)
  {
    if (r > c) std::swap(r, c);
    return r * D + c - r * (r + 1) / 2;  // upper triangle, row-major
  }
  T m_v[InternalDimension];
};

template <unsigned int D>
struct SizeOf {  // Neighborhood has a member function called Size(); name the type from outside
  typedef Size<D> Type;
};

template <typename T, unsigned int D>
class Neighborhood
{
public:
  typedef typename SizeOf<D>::Type SizeType;
  typedef typename SizeOf<D>::Type RadiusType;
  typedef Offset<D> OffsetType;
  typedef SizeValueType NeighborIndexType;
  Neighborhood() { m_Radius.Fill(0); m_Side.Fill(1); }
  void SetRadius(const SizeValueType r) { SizeType s; s.Fill(r); SetRadius(s); }
  void SetRadius(const SizeType& r)
  {
    m_Radius = r;
    NeighborIndexType n = 1;
    for (unsigned int d = 0; d < D; ++d) { m_Side[d] = 2 * r[d] + 1; n *= m_Side[d]; }
    m_Data.assign(n, T());
  }
  const SizeType& GetRadius() const { return m_Radius; }
  NeighborIndexType Size() const { return m_Data.size(); }
  T& operator[](NeighborIndexType i) { return m_Data[i]; }
  const T& operator[](NeighborIndexType i) const { return m_Data[i]; }
  T& operator[](const OffsetType& o) { return m_Data[Linear(o)]; }
  const T& operator[](const OffsetType& o) const { return m_Data[Linear(o)]; }
  OffsetType GetOffset(NeighborIndexType i) const
  {
    OffsetType o;
    for (unsigned int d = 0; d < D; ++d) {
      o[d] = static_cast<OffsetValueType>(i % m_Side[d]) - static_cast<OffsetValueType>(m_Radius[d]);
      i /= m_Side[d];
    }
    return o;
  }
private:
  NeighborIndexType Linear(const OffsetType& o) const
  {
    NeighborIndexType i = 0, stride = 1;
    for (unsigned int d = 0; d < D; ++d) {
      i += static_cast<NeighborIndexType>(o[d] + static_cast<OffsetValueType>(m_Radius[d])) * stride;
      stride *= m_Side[d];
    }
    return i;
  }
  SizeType m_Radius, m_Side;
  std::vector<T> m_Data;
};

// ---- Image -------------------------------------------------------------------------------------
template <typename TPixel, unsigned int VDim>
class Image : public LightObject
{
public:
  typedef Image Self;
  typedef SmartPointer<Self> Pointer;
  typedef SmartPointer<const Self> ConstPointer;
  typedef TPixel PixelType;
  typedef Index<VDim> IndexType;
  typedef Size<VDim> SizeType;
  typedef Offset<VDim> OffsetType;
  typedef ImageRegion<VDim> RegionType;
  typedef FixedVector<double, VDim> SpacingType;
  typedef FixedVector<double, VDim> PointType;
  typedef FixedVector<double, VDim * VDim> DirectionType;  // row-major, identity by default
  static const unsigned int ImageDimension = VDim;

  itkNewMacro(Self);
  itkTypeMacro(Image, LightObject);

  void SetRegions(const RegionType& r) { m_Region = r; }
  void SetRegions(const SizeType& s) { IndexType i; i.Fill(0); m_Region = RegionType(i, s); }
  void Allocate() { m_Buffer.assign(m_Region.GetNumberOfPixels(), TPixel()); }
  void FillBuffer(const TPixel& v) { std::fill(m_Buffer.begin(), m_Buffer.end(), v); }
  const RegionType& GetLargestPossibleRegion() const { return m_Region; }
  const RegionType& GetBufferedRegion() const { return m_Region; }
  const RegionType& GetRequestedRegion() const { return m_Region; }
  const SpacingType& GetSpacing() const { return m_Spacing; }
  void SetSpacing(const SpacingType& s) { m_Spacing = s; }
  const PointType& GetOrigin() const { return m_Origin; }
  void SetOrigin(const PointType& o) { m_Origin = o; }
  const DirectionType& GetDirection() const { return m_Direction; }
  void SetDirection(const DirectionType& d) { m_Direction = d; }
  TPixel& GetPixel(const IndexType& i) { return m_Buffer[ComputeOffset(i)]; }
  const TPixel& GetPixel(const IndexType& i) const { return m_Buffer[ComputeOffset(i)]; }
  void SetPixel(const IndexType& i, const TPixel& v) { m_Buffer[ComputeOffset(i)] = v; }
  TPixel* GetBufferPointer() { return m_Buffer.data(); }
  const TPixel* GetBufferPointer() const { return m_Buffer.data(); }
  SizeValueType ComputeOffset(const IndexType& i) const
  {
    SizeValueType o = 0, stride = 1;
    for (unsigned int d = 0; d < VDim; ++d) {
      o += static_cast<SizeValueType>(i[d] - m_Region.GetIndex(d)) * stride;
      stride *= m_Region.GetSize(d);
    }
    return o;
  }
  void Graft(const Self* o) { m_Region = o->m_Region; m_Spacing = o->m_Spacing; m_Origin = o->m_Origin; m_Direction = o->m_Direction; m_Buffer = o->m_Buffer; }

protected:
  Image()
  {
    m_Spacing.Fill(1.0);
    m_Origin.Fill(0.0);
    m_Direction.Fill(0.0);
    for (unsigned int d = 0; d < VDim; ++d) m_Direction[d * VDim + d] = 1.0;
  }
  virtual ~Image() {}

private:
  RegionType m_Region;
  SpacingType m_Spacing;
  PointType m_Origin;
  DirectionType m_Direction;
  std::vector<TPixel> m_Buffer;
};

template <unsigned int D>
std::ostream& operator<<(std::ostream& os, const ImageRegion<D>& r)
{
  os << "ImageRegion: index [";
  for (unsigned int d = 0; d < D; ++d) os << (d ? ", " : "") << r.GetIndex(d);
  os << "] size [";
  for (unsigned int d = 0; d < D; ++d) os << (d ? ", " : "") << r.GetSize(d);
  return os << "]";
}
template <typename TPixel, unsigned int VDim>
std::ostream& operator<<(std::ostream& os, const Image<TPixel, VDim>& img)
{
  os << "Image (stand-in ITK): " << img.GetLargestPossibleRegion() << " spacing [";
  for (unsigned int d = 0; d < VDim; ++d) os << (d ? ", " : "") << img.GetSpacing()[d];
  return os << "]";
}
template <typename TPixel, unsigned int VDim>
const unsigned int Image<TPixel, VDim>::ImageDimension;

// ---- region iterators (raster order over a sub-region, x fastest) -------------------------------
template <typename TImage, bool IsConst>
class RegionWalker
{
public:
  typedef typename TImage::IndexType IndexType;
  typedef typename TImage::RegionType RegionType;
  typedef typename TImage::PixelType PixelType;
  typedef typename std::conditional<IsConst, const TImage, TImage>::type ImageT;
  RegionWalker(ImageT* img, const RegionType& r) : m_Image(img), m_Region(r) { GoToBegin(); }
  void GoToBegin()
  {
    m_Index = m_Region.GetIndex();
    m_End = m_Region.GetNumberOfPixels() == 0;
  }
  bool IsAtEnd() const { return m_End; }
  void operator++()
  {
    for (unsigned int d = 0; d < TImage::ImageDimension; ++d) {
      if (++m_Index[d] < m_Region.GetIndex(d) + static_cast<IndexValueType>(m_Region.GetSize(d))) return;
      m_Index[d] = m_Region.GetIndex(d);
    }
    m_End = true;
  }
  const IndexType& GetIndex() const { return m_Index; }
  PixelType Get() const { return m_Image->GetPixel(m_Index); }
protected:
  ImageT* m_Image;
  RegionType m_Region;
  IndexType m_Index;
  bool m_End;
};

template <typename TImage>
class ImageRegionConstIterator : public RegionWalker<TImage, true>
{
public:
  typedef RegionWalker<TImage, true> Base;
  ImageRegionConstIterator(const TImage* img, const typename TImage::RegionType& r) : Base(img, r) {}
  const typename TImage::PixelType& Value() const { return this->m_Image->GetPixel(this->m_Index); }
};
template <typename TImage>
class ImageRegionIterator : public RegionWalker<TImage, false>
{
public:
  typedef RegionWalker<TImage, false> Base;
  ImageRegionIterator(TImage* img, const typename TImage::RegionType& r) : Base(img, r) {}
  typename TImage::PixelType& Value() { return this->m_Image->GetPixel(this->m_Index); }
  void Set(const typename TImage::PixelType& v) { this->m_Image->GetPixel(this->m_Index) = v; }
};
template <typename TImage>
class ImageRegionConstIteratorWithIndex : public ImageRegionConstIterator<TImage>
{
public:
  ImageRegionConstIteratorWithIndex(const TImage* img, const typename TImage::RegionType& r) : ImageRegionConstIterator<TImage>(img, r) {}
};
template <typename TImage>
class ImageRegionIteratorWithIndex : public ImageRegionIterator<TImage>
{
public:
  ImageRegionIteratorWithIndex(TImage* img, const typename TImage::RegionType& r) : ImageRegionIterator<TImage>(img, r) {}
};

// ---- neighbourhood iterators: offsets are resolved against the live image -------------------------
template <typename TImage>
class ConstNeighborhoodIterator : public RegionWalker<TImage, true>
{
public:
  typedef typename TImage::OffsetType OffsetType;
  typedef typename TImage::SizeType RadiusType;
  ConstNeighborhoodIterator(const RadiusType&, const TImage* img, const typename TImage::RegionType& r) : RegionWalker<TImage, true>(img, r) {}
  typename TImage::PixelType GetPixel(const OffsetType& o) const { return this->m_Image->GetPixel(this->m_Index + o); }
  typename TImage::PixelType GetCenterPixel() const { return this->m_Image->GetPixel(this->m_Index); }
};
template <typename TImage>
class NeighborhoodIterator : public RegionWalker<TImage, false>
{
public:
  typedef typename TImage::OffsetType OffsetType;
  typedef typename TImage::SizeType RadiusType;
  NeighborhoodIterator(const RadiusType&, TImage* img, const typename TImage::RegionType& r) : RegionWalker<TImage, false>(img, r) {}
  typename TImage::PixelType GetPixel(const OffsetType& o) const { return this->m_Image->GetPixel(this->m_Index + o); }
  typename TImage::PixelType GetCenterPixel() const { return this->m_Image->GetPixel(this->m_Index); }
  void SetCenterPixel(const typename TImage::PixelType& v) { this->m_Image->GetPixel(this->m_Index) = v; }
};

namespace NeighborhoodAlgorithm
{
// Splits `region` into the part whose radius-neighbourhood stays inside the image buffer (first list entry)
// and non-overlapping boundary faces (remaining entries), dimension by dimension.
template <typename TImage>
struct ImageBoundaryFacesCalculator {
  typedef typename TImage::RegionType RegionType;
  typedef typename TImage::SizeType RadiusType;
  typedef typename TImage::IndexType IndexType;
  typedef typename TImage::SizeType SizeType;
  typedef std::list<RegionType> FaceListType;
  FaceListType operator()(const TImage* img, RegionType region, RadiusType radius) const
  {
    const unsigned int D = TImage::ImageDimension;
    const RegionType& buf = img->GetBufferedRegion();
    FaceListType faces;
    IndexType rest_i = region.GetIndex();
    SizeType rest_s = region.GetSize();
    for (unsigned int d = 0; d < D; ++d) {
      const IndexValueType b0 = buf.GetIndex(d), b1 = b0 + static_cast<IndexValueType>(buf.GetSize(d));
      const IndexValueType lo = rest_i[d], hi = lo + static_cast<IndexValueType>(rest_s[d]);
      const IndexValueType in_lo = std::min(std::max(b0 + static_cast<IndexValueType>(radius[d]), lo), hi);
      const IndexValueType in_hi = std::min(std::max(b1 - static_cast<IndexValueType>(radius[d]), in_lo), hi);
      if (in_lo > lo) {
        IndexType fi = rest_i; SizeType fs = rest_s;
        fi[d] = lo; fs[d] = static_cast<SizeValueType>(in_lo - lo);
        faces.push_back(RegionType(fi, fs));
      }
      if (in_hi < hi) {
        IndexType fi = rest_i; SizeType fs = rest_s;
        fi[d] = in_hi; fs[d] = static_cast<SizeValueType>(hi - in_hi);
        faces.push_back(RegionType(fi, fs));
      }
      rest_i[d] = in_lo;
      rest_s[d] = static_cast<SizeValueType>(in_hi - in_lo);
    }
    bool empty = false;
    for (unsigned int d = 0; d < D; ++d) empty = empty || rest_s[d] == 0;
    if (empty) rest_s.Fill(0);
    faces.push_front(RegionType(rest_i, rest_s));
    return faces;
  }
};
}  // namespace NeighborhoodAlgorithm

template <typename TImage>
class ImageDuplicator : public LightObject
{
public:
  typedef ImageDuplicator Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro(Self);
  void SetInputImage(const TImage* img) { m_Input = img; }
  void Update()
  {
    m_Output = TImage::New();
    m_Output->Graft(m_Input);
  }
  TImage* GetOutput() { return m_Output.GetPointer(); }
protected:
  ImageDuplicator() : m_Input(nullptr) {}
private:
  const TImage* m_Input;
  typename TImage::Pointer m_Output;
};

template <typename TInputImage, typename TOutputImage>
class ImageToImageFilter : public LightObject
{
public:
  typedef ImageToImageFilter Self;
  typedef SmartPointer<Self> Pointer;
  void SetInput(const TInputImage* in) { m_Input = in; }
  const TInputImage* GetInput() const { return m_Input; }
  TOutputImage* GetOutput() { return m_Output.GetPointer(); }
  void Update() { this->GenerateData(); }
protected:
  ImageToImageFilter() : m_Input(nullptr) {}
  virtual ~ImageToImageFilter() {}
  virtual void GenerateData() = 0;
  void AllocateOutputs() {}
  void GraftOutput(TOutputImage* out) { m_Output = out; }
private:
  const TInputImage* m_Input;
  typename TOutputImage::Pointer m_Output;
};
}  // namespace itk

// ---- vnl ----------------------------------------------------------------------------------------
template <typename T>
class vnl_vector
{
public:
  vnl_vector() {}
  explicit vnl_vector(size_t n) : m_d(n) {}
  vnl_vector(size_t n, const T& v) : m_d(n, v) {}
  size_t size() const { return m_d.size(); }
  T& operator()(size_t i) { return m_d[i]; }
  const T& operator()(size_t i) const { return m_d[i]; }
  T& operator[](size_t i) { return m_d[i]; }
  const T& operator[](size_t i) const { return m_d[i]; }
private:
  std::vector<T> m_d;
};

template <typename T>
class vnl_sparse_matrix
{
public:
  vnl_sparse_matrix(unsigned int r, unsigned int c) : m_rows(r), m_cols(c), m_row(r) {}
  T& operator()(unsigned int r, unsigned int c) { return m_row[r][c]; }
  unsigned int rows() const { return m_rows; }
  unsigned int cols() const { return m_cols; }
  const std::map<unsigned int, T>& row(unsigned int r) const { return m_row[r]; }
private:
  unsigned int m_rows, m_cols;
  std::vector<std::map<unsigned int, T> > m_row;
};

// Exact solve standing in for VXL's sparse LU: dense LU with partial pivoting, factored once.
class vnl_sparse_lu
{
public:
  explicit vnl_sparse_lu(const vnl_sparse_matrix<double>& M) : m_n(M.rows()), m_a(static_cast<size_t>(M.rows()) * M.rows(), 0.0), m_piv(M.rows())
  {
    const size_t n = m_n;
    for (unsigned int r = 0; r < m_n; ++r)
      for (std::map<unsigned int, double>::const_iterator it = M.row(r).begin(); it != M.row(r).end(); ++it) m_a[r * n + it->first] = it->second;
    for (size_t k = 0; k < n; ++k) {
      size_t p = k;
      double best = std::fabs(m_a[k * n + k]);
      for (size_t i = k + 1; i < n; ++i)
        if (std::fabs(m_a[i * n + k]) > best) { best = std::fabs(m_a[i * n + k]); p = i; }
      if (best == 0.0) throw std::runtime_error("vnl_sparse_lu: singular matrix");
      m_piv[k] = p;
      if (p != k)
        for (size_t j = 0; j < n; ++j) std::swap(m_a[k * n + j], m_a[p * n + j]);
      const double inv = 1.0 / m_a[k * n + k];
      for (size_t i = k + 1; i < n; ++i) {
        double m = m_a[i * n + k];
        if (m == 0.0) continue;
        m *= inv;
        m_a[i * n + k] = m;
        for (size_t j = k + 1; j < n; ++j) m_a[i * n + j] -= m * m_a[k * n + j];
      }
    }
  }
  vnl_vector<double> solve(const vnl_vector<double>& b) const
  {
    const size_t n = m_n;
    vnl_vector<double> x(n);
    for (size_t i = 0; i < n; ++i) x(i) = b(i);
    for (size_t k = 0; k < n; ++k)
      if (m_piv[k] != k) std::swap(x(k), x(m_piv[k]));
    for (size_t i = 1; i < n; ++i) {
      double s = x(i);
      for (size_t j = 0; j < i; ++j) s -= m_a[i * n + j] * x(j);
      x(i) = s;
    }
    for (size_t ii = n; ii-- > 0;) {
      double s = x(ii);
      for (size_t j = ii + 1; j < n; ++j) s -= m_a[ii * n + j] * x(j);
      x(ii) = s / m_a[ii * n + ii];
    }
    return x;
  }
private:
  unsigned int m_n;
  std::vector<double> m_a;
  std::vector<size_t> m_piv;
};

#endif  // MINI_ITK_H
