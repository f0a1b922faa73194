// Common enums and helpers shared by the host runtime, the CPU backend and the
// CUDA backend.  Torch-free on purpose: only csrc/api/ includes torch headers.
//
// Parity notes (reference = helmholtz-analytics/mpi4torch):
//   * ReduceOp values 0..11 mirror the integer constants exported by the
//     reference (csrc/extension.cpp:204-218, :1424-1435).
//   * DType covers the reference's dtype map (csrc/extension.cpp:106-129:
//     u8,i8,i16,i32,i64,f32,f64) plus bf16/f16/bool which the reference rejects.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>

namespace m4t {

enum class DType : int32_t {
  U8 = 0, I8 = 1, I16 = 2, I32 = 3, I64 = 4, F32 = 5, F64 = 6, BF16 = 7, F16 = 8, BOOL = 9,
  kCount = 10
};

inline int64_t dtype_size(DType d) {
  switch (d) {
    case DType::U8: case DType::I8: case DType::BOOL: return 1;
    case DType::I16: case DType::BF16: case DType::F16: return 2;
    case DType::I32: case DType::F32: return 4;
    case DType::I64: case DType::F64: return 8;
    default: return 0;
  }
}

inline const char* dtype_name(DType d) {
  static const char* names[] = {"u8", "i8", "i16", "i32", "i64", "f32", "f64", "bf16", "f16", "bool"};
  int i = static_cast<int>(d);
  return (i >= 0 && i < static_cast<int>(DType::kCount)) ? names[i] : "?";
}

inline bool dtype_is_float(DType d) {
  return d == DType::F32 || d == DType::F64 || d == DType::BF16 || d == DType::F16;
}

enum class ReduceOp : int32_t {
  MAX = 0, MIN = 1, SUM = 2, PROD = 3, LAND = 4, BAND = 5, LOR = 6, BOR = 7, LXOR = 8, BXOR = 9,
  MINLOC = 10, MAXLOC = 11, kCount = 12
};

inline const char* op_name(ReduceOp op) {
  static const char* names[] = {"MAX", "MIN", "SUM", "PROD", "LAND", "BAND", "LOR", "BOR",
                                "LXOR", "BXOR", "MINLOC", "MAXLOC"};
  int i = static_cast<int>(op);
  return (i >= 0 && i < static_cast<int>(ReduceOp::kCount)) ? names[i] : "?";
}

// Validates an (op, dtype) pair.  MINLOC/MAXLOC exist as constants only: the
// reference can never run them either because no pair datatype is ever built
// (csrc/extension.cpp:106-129 yields scalar types only).
inline void check_op_dtype(ReduceOp op, DType dt) {
  int o = static_cast<int>(op);
  if (o < 0 || o >= static_cast<int>(ReduceOp::kCount))
    throw std::invalid_argument("mpi4torch_b200: Collective operation not supported!");
  if (op == ReduceOp::MINLOC || op == ReduceOp::MAXLOC)
    throw std::invalid_argument(
        "mpi4torch_b200: MPI_MINLOC/MPI_MAXLOC need a (value,index) pair datatype, which is not "
        "representable as a torch dtype");
  if ((op == ReduceOp::BAND || op == ReduceOp::BOR || op == ReduceOp::BXOR) && dtype_is_float(dt))
    throw std::invalid_argument(std::string("mpi4torch_b200: bitwise reduction ") + op_name(op) +
                                " is undefined for floating dtype " + dtype_name(dt));
}

// Fused epilogue applied by every reducing collective:
//   out = cast_out( acc_in + scale * reduce(...) )
// `scale` is applied in the accumulation precision (fp32 for 16-bit floats).
struct Epilogue {
  double scale = 1.0;
  bool has_scale = false;
  const void* accumulate = nullptr;  // optional tensor (out dtype) added to the result
};

#define M4T_CHECK(cond, ...)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      std::ostringstream m4t_oss_;                                                  \
      m4t_oss_ << "mpi4torch_b200: " << __VA_ARGS__ << " (" << __FILE__ << ":" << __LINE__ << ")"; \
      throw std::runtime_error(m4t_oss_.str());                                     \
    }                                                                               \
  } while (0)

inline int64_t env_i64(const char* name, int64_t dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  return std::strtoll(v, nullptr, 10);
}

inline bool debug_enabled() {
  static int v = static_cast<int>(env_i64("M4T_DEBUG", 0));
  return v != 0;
}

#define M4T_LOG(...)                                 \
  do {                                               \
    if (::m4t::debug_enabled()) {                    \
      std::fprintf(stderr, "[m4t] " __VA_ARGS__);    \
      std::fprintf(stderr, "\n");                    \
    }                                                \
  } while (0)

}  // namespace m4t
