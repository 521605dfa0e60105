// Counter slots of the dynamic tile scheduler (cube_sched.cuh).
#include <atomic>
#include <cstdlib>
#include "cube_sched.cuh"

namespace sched {

__device__ Slot g_slots[kSlots];          // zero-initialised; every kernel re-arms its slot when it ends

Slot* claim_slot()
{
    static std::atomic<unsigned> seq{0};
    static Slot* base[64] = {};           // per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    Slot* b = base[dev];
    if (!b) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_slots) != cudaSuccess) return nullptr;
        b = base[dev] = static_cast<Slot*>(p);
    }
    return b + (seq.fetch_add(1, std::memory_order_relaxed) % kSlots);
}

int tail_div()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CUBE_TAIL_DIV");
        v = e ? atoi(e) : kTailDiv;
        if (v < 0) v = kTailDiv;
    }
    return v;
}

}  // namespace sched
