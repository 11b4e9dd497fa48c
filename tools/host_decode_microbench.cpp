// Probe (not part of the product): the inner loop of the AVX-512 VBMI two-bit decoder (csrc/host_expand.cpp) against a
// plain non-temporal fill, per thread count.  Modes: 2 fill, 0 decode (stream read from memory, software prefetch),
// 3 decode without the prefetch, 1 decode with the stream served from L1 (no memory reads).
//   g++ -O3 -pthread tools/host_decode_microbench.cpp -o /tmp/mb && /tmp/mb 1 && /tmp/mb 8
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
struct T { alignas(64) uint8_t idx[64], ctl[64], lut[64]; };
static T mk() { T x; for (int k=0;k<8;++k) for (int i=0;i<8;++i){ x.idx[8*k+i]=(uint8_t)(2*k+(i<3?i:0)); x.ctl[8*k+i]=(uint8_t)(2*i);} for(int i=0;i<64;++i) x.lut[i]="ACGT"[i&3]; return x; }
static T tab = mk();
// mode: 0 decode big->big NT, 1 decode small(L1) input -> big NT, 2 fill NT, 3 decode big->big NT no prefetch, 4 decode with 2 streams interleaved
__attribute__((target("avx512f,avx512bw,avx512vbmi")))
static void run(int mode, uint8_t* dst, const uint8_t* src, size_t nlines, size_t piece_lines) {
  const __m512i idx=_mm512_load_si512(tab.idx), ctl=_mm512_load_si512(tab.ctl), lut=_mm512_load_si512(tab.lut), m3=_mm512_set1_epi8(3);
  const __m512i fillv=_mm512_set1_epi8('A');
  for (size_t base=0; base<nlines; base+=piece_lines) {
    size_t n = std::min(piece_lines, nlines-base);
    uint8_t* d = dst + base*64; const uint8_t* q = (mode==1) ? src : src + base*16;
    if (mode==2) { for (size_t g=0; g<n; ++g, d+=64) _mm512_stream_si512((__m512i*)d, fillv); continue; }
    for (size_t g=0; g<n; ++g, d+=64) {
      if (mode==0) _mm_prefetch((const char*)q+1024,_MM_HINT_T0);
      const __m512i s=_mm512_castsi256_si512(_mm256_loadu_si256((const __m256i*)q));
      const __m512i rep=_mm512_permutexvar_epi8(idx,s);
      const __m512i code=_mm512_and_si512(_mm512_multishift_epi64_epi8(ctl,rep),m3);
      _mm512_stream_si512((__m512i*)d,_mm512_shuffle_epi8(lut,code));
      q += (mode==1) ? ((g&63)==63 ? -1008 : 16) : 16;
    }
  }
  _mm_sfence();
}
int main(int argc,char**argv){
  int threads = argc>1?atoi(argv[1]):1; size_t gb = 4;
  size_t out_bytes = gb<<30, nlines=out_bytes/64;
  uint8_t* out=(uint8_t*)aligned_alloc(4096,out_bytes); uint8_t* in=(uint8_t*)aligned_alloc(4096,out_bytes/4+4096);
  memset(out,1,out_bytes); for(size_t i=0;i<out_bytes/4+4096;++i) in[i]=(uint8_t)(i*2654435761u>>13);
  for (int mode : {2,0,3,1,2,0}) {
    auto t0=std::chrono::steady_clock::now();
    std::vector<std::thread> th; size_t per=nlines/threads;
    for(int t=0;t<threads;++t) th.emplace_back([&,t]{ run(mode,out+t*per*64,in+t*per*16,per,320); });
    for(auto&x:th)x.join();
    double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
    printf("threads %d mode %d: %.1f GB/s (%.2f per thread)\n",threads,mode,out_bytes/dt/1e9,out_bytes/dt/1e9/threads);
  }
}
