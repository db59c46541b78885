// TEST-ONLY host emulation of the infidex_b200 kernels (Ctx == one host thread). Lets the CPU test suite check the
// kernel logic against the oracle without a GPU. Never loaded by the product package.
#define IFX_EMU 1
#include "../../infidex_b200/csrc/ifx_api.inl"
