// Compile-time latent sizes served by the register-resident (thread-per-lane) kernels.
// Every other nseasons goes through the shared-memory warp-per-lane path (ci_bigd.cu).
#pragma once

#define CI_MAX_REG_S 12

#define CI_DISPATCH_S(S_runtime, CALL)                 \
  switch (S_runtime) {                                 \
    case 1: { constexpr int S_ = 1; CALL; } break;     \
    case 2: { constexpr int S_ = 2; CALL; } break;     \
    case 3: { constexpr int S_ = 3; CALL; } break;     \
    case 4: { constexpr int S_ = 4; CALL; } break;     \
    case 5: { constexpr int S_ = 5; CALL; } break;     \
    case 6: { constexpr int S_ = 6; CALL; } break;     \
    case 7: { constexpr int S_ = 7; CALL; } break;     \
    case 12: { constexpr int S_ = 12; CALL; } break;   \
    default: return -3; /* CI_ERR_UNSUPPORTED */       \
  }

// nseasons served by the register-resident kernels (must match the cases above)
static inline bool ci_reg_set(int S) { return (S >= 1 && S <= 7) || S == 12; }
