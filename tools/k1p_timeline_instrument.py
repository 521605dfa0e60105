"""Patch a COPY of csrc/scramble.cu with the per-warp %globaltimer stamps tools/k1p_timeline.py reads (never the
product tree):  cp -r rubiks_cube_solver_b200/csrc include <scratch>; python tools/k1p_timeline_instrument.py
<scratch>/csrc/scramble.cu; python <scratch>/csrc/build.py"""
import sys
p=sys.argv[1]; s=open(p).read()
def rep(a,b):
    global s
    assert a in s, a[:60]
    s=s.replace(a,b,1)
rep('''namespace {

#ifndef CUBE_SCRAMBLE_MIN_BLOCKS''','''__device__ unsigned long long g_tl[4][148][32][6];
extern "C" int cube_debug_timeline(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_tl, sizeof(g_tl)); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

namespace {

#ifndef CUBE_SCRAMBLE_MIN_BLOCKS''')
rep('''    asm volatile("griddepcontrol.launch_dependents;");
    pair_table_fill<SIZE>(s_ptbl, tid, blockDim.x);''','''    const unsigned long long t_entry = gtime();
    asm volatile("griddepcontrol.launch_dependents;");
    pair_table_fill<SIZE>(s_ptbl, tid, blockDim.x);''')
rep('''    asm volatile("griddepcontrol.wait;" ::: "memory");                // everything earlier in the stream is complete
''','''    const unsigned long long t_fill = gtime();
    asm volatile("griddepcontrol.wait;" ::: "memory");                // everything earlier in the stream is complete
    const unsigned long long t_wait = gtime();
    unsigned long long t_first = 0;
''')
rep('''        bulk::mbar_wait(&s_bar[buf], (uint32_t)(it >> 1) & 1u);

        CubieState st[NS];''','''        bulk::mbar_wait(&s_bar[buf], (uint32_t)(it >> 1) & 1u);
        if (it == 0) t_first = gtime();

        CubieState st[NS];''')
rep('''    if (lane == 0) bulk::wait_read_all();                             // shared memory must outlive the copies' reads
    __syncthreads();
    sched::release(slot);''','''    if (lane == 0) bulk::wait_read_all();                             // shared memory must outlive the copies' reads
    if (lane == 0 && blockIdx.x < 148) {
        unsigned long long* r = g_tl[(reinterpret_cast<uintptr_t>(slot) / sizeof(sched::Slot)) & 3][blockIdx.x][warp];
        unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        r[0] = t_entry; r[1] = t_fill; r[2] = t_wait; r[3] = t_first; r[4] = gtime(); r[5] = smid;
    }
    __syncthreads();
    sched::release(slot);''')
open(p,'w').write(s)
