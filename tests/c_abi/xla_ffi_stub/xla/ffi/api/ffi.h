// STAND-IN for XLA's xla/ffi/api/ffi.h -- test infrastructure only, NOT the real header.
//
// JAX / XLA cannot be installed in this image, so jaxovercooked_b200/csrc/xla_ffi_shim.cc can never be compiled against
// the real FFI here.  This stub models just enough of the public binding DSL (Ffi::Bind().Ctx<>().Attr<>().Arg<>().Ret<>(),
// Buffer<dtype>, Result<Buffer<dtype>>, Error, XLA_FFI_DEFINE_HANDLER_SYMBOL) for a COMPILE-ONLY check
// (tests/test_abi_and_host.py::test_xla_ffi_shim_compiles_against_the_stub): the shim is valid C++, every handler's
// parameter list matches its binding position by position (the macro static_asserts it), and every call into
// include/oc_b200.h type-checks.  It says nothing about run-time behaviour under XLA.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { PRED, S8, S32, U8, U32, F32 };
template <DataType D> struct NativeType;
template <> struct NativeType<PRED> { using type = bool; };
template <> struct NativeType<S8> { using type = int8_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<F32> { using type = float; };

template <typename T>
struct Span {
    const T *ptr = nullptr;
    size_t n = 0;
    size_t size() const { return n; }
    const T &operator[](size_t i) const { return ptr[i]; }
};

template <DataType D>
struct Buffer {
    using T = typename NativeType<D>::type;
    T *data = nullptr;
    T *typed_data() const { return data; }
    Span<int64_t> dimensions() const { return {}; }
    size_t element_count() const { return 0; }
};

template <typename B>
struct Result {
    B value;
    B *operator->() { return &value; }
    B &operator*() { return value; }
};
template <DataType D> using ResultBuffer = Result<Buffer<D>>;

struct Error {
    static Error Success() { return {}; }
    static Error Internal(const std::string &) { return {}; }
};

template <typename S> struct PlatformStream { using type = S; };

template <typename... Ts>
struct Binding {
    template <typename C> Binding<Ts..., typename C::type> Ctx() const { return {}; }
    template <typename A> Binding<Ts..., A> Arg() const { return {}; }
    template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
    template <typename A> Binding<Ts..., A> Attr(const char *) const { return {}; }
    template <typename F> static constexpr bool Matches() { return std::is_invocable_r<Error, F, Ts...>::value; }
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an exported XLA_FFI_Handler; the stand-in checks that the implementation is callable with exactly
// the bound context / attribute / argument / result types, in binding order
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                                              \
    static_assert(decltype(binding)::template Matches<decltype(&impl)>(), #sym ": handler signature does not match its binding"); \
    extern "C" int sym() { return 0; }
