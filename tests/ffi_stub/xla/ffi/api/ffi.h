// STAND-IN for jaxlib's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE ONLY.
// jaxlib is not installed in this image, so mava_b200/csrc/xla_ffi_shim.cc cannot be compiled
// against the real header here.  This file declares just enough of the documented FFI C++ API
// (Buffer / Result / Span / Error / the binding DSL / the handler macro) for the compiler to
// type-check the shim's handler signatures against their bindings and, above all, its calls into
// include/mava_b200.h.  It proves nothing about XLA itself.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;

namespace xla {
namespace ffi {

enum DataType { U8, S8, U32, S32, F32 };
template <DataType T> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S8> { using type = int8_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<F32> { using type = float; };

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType T>
struct Buffer {
  using Native = typename NativeOf<T>::type;
  Native* data = nullptr;
  Span<int64_t> dims;
  Native* typed_data() const { return data; }
  size_t element_count() const { return 0; }
  Span<int64_t> dimensions() const { return dims; }
};

template <typename B>
struct Result {
  B value;
  B* operator->() { return &value; }
};
template <DataType T>
using ResultBuffer = Result<Buffer<T>>;

struct Error {
  static Error Success() { return {}; }
  static Error Internal(const std::string&) { return {}; }
  static Error InvalidArgument(const std::string&) { return {}; }
};

template <typename T> struct PlatformStream {};

// The binding DSL accumulates the C++ argument types the handler must accept.
template <typename... Ts>
struct Binding {
  template <typename C> auto Ctx() const;
  template <typename B> Binding<Ts..., B> Arg() const { return {}; }
  template <typename B> Binding<Ts..., Result<B>> Ret() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
};
template <typename C> struct CtxType;
template <typename S> struct CtxType<PlatformStream<S>> { using type = S; };
template <typename... Ts>
template <typename C>
auto Binding<Ts...>::Ctx() const { return Binding<Ts..., typename CtxType<C>::type>{}; }

struct Ffi {
  static Binding<> Bind() { return {}; }
};

template <typename... Ts, typename Fn>
constexpr bool CheckHandler(Binding<Ts...>, Fn) {
  static_assert(std::is_invocable_r<Error, Fn, Ts...>::value,
                "handler signature does not match its binding");
  return true;
}

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, fn, binding)                     \
  static const bool sym##_checked = ::xla::ffi::CheckHandler(binding, fn);  \
  extern "C" XLA_FFI_Error* sym(XLA_FFI_CallFrame*) { return nullptr; }
