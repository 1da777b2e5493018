// TEST INFRASTRUCTURE — a minimal stand-in for jaxlib's xla/ffi/api/ffi.h (absent from this image), just enough of the
// binding DSL to COMPILE kalman_jax_b200/csrc/kjx_xla_ffi.cc and to type-check every handler against its binding:
// Bind().Ctx<>().Attr<>().Arg<>().Ret<>() accumulates the C++ argument types and XLA_FFI_DEFINE_HANDLER_SYMBOL asserts the
// handler is invocable with exactly those.  Shapes follow the public XLA FFI API (Buffer / Result / Span / Error); no
// behaviour is emulated.  Used by tests/test_xla_ffi_stub.py with -fsyntax-only.
#ifndef KJX_TEST_XLA_FFI_STUB_H_
#define KJX_TEST_XLA_FFI_STUB_H_
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <type_traits>

namespace xla { namespace ffi {

enum DataType { U8, F32, F64, S64 };
template <DataType dt> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<S64> { using type = int64_t; };

template <typename T> struct Span {
  const T* data() const { return ptr; }
  size_t size() const { return n; }
  const T* ptr = nullptr; size_t n = 0;
};

template <DataType dt> struct Buffer {
  using Native = typename NativeOf<dt>::type;
  Native* typed_data() const { return static_cast<Native*>(ptr); }
  void* untyped_data() const { return ptr; }
  size_t element_count() const { return count; }
  size_t size_bytes() const { return count * sizeof(Native); }
  Span<int64_t> dimensions() const { return {}; }
  void* ptr = nullptr; size_t count = 0;
};
template <typename T> struct Result {
  T* operator->() { return &value; }
  T& operator*() { return value; }
  T value;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kOk, kInvalidArgument, kUnimplemented, kInternal };
struct Error {
  Error() = default;
  Error(ErrorCode c, std::string m) : code(c), message(std::move(m)) {}
  static Error Success() { return Error(); }
  ErrorCode code = ErrorCode::kOk; std::string message;
};

template <typename S> struct PlatformStream { using type = S; };

template <typename... Ts> struct Binding {
  template <typename C> Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(std::string_view) const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename F> static constexpr bool Accepts() { return std::is_invocable_r<Error, F, Ts...>::value; }
};
struct Ffi { static Binding<> Bind() { return {}; } };

}}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                               \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(),                                    \
                #impl " does not take the arguments its binding declares");                                \
  extern "C" void* symbol(void*) { return nullptr; }

#endif
