// pic_xla_ffi.cc -- XLA FFI custom-call handlers over the C ABI of include/pic_b200.h: the north star's "C ABI exposed as JAX FFI
// custom calls".  Each handler receives device buffers + the CUDA stream XLA runs the call on + the PicParams struct as an
// opaque byte attribute, and forwards to the export that replaces the reference call site named beside it.
//
// Compile-gated: this image has no jax / jaxlib, so `xla/ffi/api/ffi.h` is not on disk and the file compiles to an empty
// translation unit here (tests/test_abi.py checks exactly that); where jaxlib is installed it is built with
//     g++ -std=c++17 -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//         -I/usr/local/cuda/include pypic3d_b200/csrc/pic_xla_ffi.cc -Lpypic3d_b200 -lpic_b200 -o pypic3d_b200/libpic_xla_ffi.so
// and registered from Python:
//     jax.ffi.register_ffi_target("pic_update_B", jax.ffi.pycapsule(lib.PicUpdateB), platform="CUDA")
//     jax.ffi.ffi_call("pic_update_B", out_types, input_output_aliases={3: 0, 4: 1, 5: 2})(Ex, Ey, Ez, Bx, By, Bz, params=...)
// Functional contract of the reference (inputs are never mutated): in-place exports take their target through
// input_output_aliases, so XLA hands the handler a result buffer that already holds the input values (or a copy of them).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define PIC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef PIC_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "pic_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using Buf = ffi::AnyBuffer;
using Res = ffi::Result<ffi::AnyBuffer>;
using Bytes = ffi::Span<const uint8_t>;

inline const PicParams* params_of(Bytes b, ffi::Error* err) {
    if (b.size() != (size_t)pic_params_size()) {
        *err = ffi::Error::InvalidArgument("params: expected the bytes of a PicParams struct (pic_params_size())");
        return nullptr;
    }
    return reinterpret_cast<const PicParams*>(b.data());
}
inline ffi::Error status(int rc, const char* what) { return rc ? ffi::Error::Internal(what) : ffi::Error::Success(); }

// solvers/first_order_yee.py:116-142 update_B (half step): B aliased to the results
ffi::Error UpdateBImpl(cudaStream_t st, Buf ex, Buf ey, Buf ez, Res bx, Res by, Res bz, Bytes params) {
    ffi::Error e = ffi::Error::Success();
    const PicParams* p = params_of(params, &e);
    if (!p) return e;
    const void* E[3] = {ex.untyped_data(), ey.untyped_data(), ez.untyped_data()};
    void* B[3] = {bx->untyped_data(), by->untyped_data(), bz->untyped_data()};
    return status(pic_update_B(p, B, E, st), "pic_update_B");
}
// solvers/first_order_yee.py:42-72 update_E: E aliased to the results
ffi::Error UpdateEImpl(cudaStream_t st, Buf bx, Buf by, Buf bz, Buf jx, Buf jy, Buf jz, Res ex, Res ey, Res ez, Bytes params) {
    ffi::Error e = ffi::Error::Success();
    const PicParams* p = params_of(params, &e);
    if (!p) return e;
    const void* B[3] = {bx.untyped_data(), by.untyped_data(), bz.untyped_data()};
    const void* J[3] = {jx.untyped_data(), jy.untyped_data(), jz.untyped_data()};
    void* E[3] = {ex->untyped_data(), ey->untyped_data(), ez->untyped_data()};
    return status(pic_update_E(p, E, B, J, st), "pic_update_E");
}
// evolve.py:88-96 in one pass: B(half) -> E -> B(half); outputs are separate arrays
ffi::Error YeeFusedImpl(cudaStream_t st, Buf ex, Buf ey, Buf ez, Buf bx, Buf by, Buf bz, Buf jx, Buf jy, Buf jz, Res ex2, Res ey2,
                        Res ez2, Res bx2, Res by2, Res bz2, Bytes params) {
    ffi::Error e = ffi::Error::Success();
    const PicParams* p = params_of(params, &e);
    if (!p) return e;
    const void* E[3] = {ex.untyped_data(), ey.untyped_data(), ez.untyped_data()};
    const void* B[3] = {bx.untyped_data(), by.untyped_data(), bz.untyped_data()};
    const void* J[3] = {jx.untyped_data(), jy.untyped_data(), jz.untyped_data()};
    void* E2[3] = {ex2->untyped_data(), ey2->untyped_data(), ez2->untyped_data()};
    void* B2[3] = {bx2->untyped_data(), by2->untyped_data(), bz2->untyped_data()};
    return status(pic_yee_fused(p, E, B, J, E2, B2, st), "pic_yee_fused");
}
// pusher/particle_push.py:13-175 particle_push: u_out is a fresh result
ffi::Error PushImpl(cudaStream_t st, Buf x, Buf u, Buf active, Buf ex, Buf ey, Buf ez, Buf bx, Buf by, Buf bz, Res u_out, int64_t cap,
                    Bytes params) {
    ffi::Error e = ffi::Error::Success();
    const PicParams* p = params_of(params, &e);
    if (!p) return e;
    const void* E[3] = {ex.untyped_data(), ey.untyped_data(), ez.untyped_data()};
    const void* B[3] = {bx.untyped_data(), by.untyped_data(), bz.untyped_data()};
    return status(pic_push(p, x.untyped_data(), u.untyped_data(), u_out->untyped_data(), (const uint8_t*)active.untyped_data(), cap, E, B, st),
                  "pic_push");
}
// deposition/Esirkepov.py:49-504 Esirkepov_current: J aliased to zero-initialised results
ffi::Error EsirkepovImpl(cudaStream_t st, Buf x, Buf u, Buf active, Res jx, Res jy, Res jz, int64_t cap, Bytes params) {
    ffi::Error e = ffi::Error::Success();
    const PicParams* p = params_of(params, &e);
    if (!p) return e;
    void* J[3] = {jx->untyped_data(), jy->untyped_data(), jz->untyped_data()};
    return status(pic_deposit_esirkepov(p, x.untyped_data(), u.untyped_data(), (const uint8_t*)active.untyped_data(), cap, J, st),
                  "pic_deposit_esirkepov");
}

}  // namespace

#define PIC_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(PicUpdateB, UpdateBImpl,
                              PIC_STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<Bytes>("params"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(PicUpdateE, UpdateEImpl,
                              PIC_STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>()
                                  .Attr<Bytes>("params"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(PicYeeFused, YeeFusedImpl,
                              PIC_STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>()
                                  .Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<Bytes>("params"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(PicPush, PushImpl,
                              PIC_STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>()
                                  .Ret<Buf>().Attr<int64_t>("cap").Attr<Bytes>("params"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(PicEsirkepov, EsirkepovImpl,
                              PIC_STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int64_t>("cap")
                                  .Attr<Bytes>("params"));
#undef PIC_STREAM

#else  // no XLA FFI headers in this environment: nothing to build (the C ABI itself is the boundary, bound with ctypes)
extern "C" int pic_xla_ffi_available(void) { return 0; }
#endif
