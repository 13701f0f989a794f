// MOCK of the XLA FFI C++ API surface (xla/ffi/api/ffi.h, header-only, shipped in jaxlib >= 0.4.31).
// TEST INFRASTRUCTURE ONLY: jaxlib is not installable in this image, so tests/test_abi.py compiles
// flowmc_b200/csrc/flowmc_xla_ffi.cc against this stand-in to keep the shim's calls into the C ABI of
// include/flowmc_b200.h type-checked (argument order, pointer types, struct fields).  It declares just the names the
// shim uses, with the shapes they have in the real header; it binds nothing and runs nothing.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>

namespace xla {
namespace ffi {

enum DataType { F32, U32, S32 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<S32> { using type = int32_t; };

template <class T>
class Span {
 public:
  Span() : p_(nullptr), n_(0) {}
  Span(T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  T& operator[](size_t i) const { return p_[i]; }

 private:
  T* p_;
  size_t n_;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeType<dtype>::type;
  T* typed_data() const { return nullptr; }
  Span<const int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
};

class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  Span<const int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
};

template <class T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument, kResourceExhausted };

class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool failure() const { return code_ != ErrorCode::kOk; }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t, size_t = 1) { return std::nullopt; }
};

template <class S>
struct PlatformStream {};

struct Binding {
  template <class T> Binding& Ctx() { return *this; }
  template <class T> Binding& Arg() { return *this; }
  template <class T> Binding& Ret() { return *this; }
  template <class T> Binding& Attr(const char*) { return *this; }
};

struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines `extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame*)`; the mock only references the
// implementation so that it is instantiated and type-checked
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)              \
  extern "C" void* fn##_mock_symbol() {                               \
    (void)(binding);                                                  \
    return reinterpret_cast<void*>(&impl);                            \
  }
