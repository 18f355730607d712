// CPU test of the host copy pool behind the pageable-buffer path (x266_b200/csrc/hostcopy.cpp): every thread count and
// store kind copies exactly, misaligned heads/tails included, and concurrent callers share the pool.  Prints "ok" on success.
#include "hostcopy.h"
#include "chunk_claimer.h"
#include <vector>
#include <cstring>
#include <cstdio>
#include <chrono>
#include <thread>
using namespace x266;
int main(){
  size_t n = 24<<20; int bad = 0; std::vector<char> a(n), b(n), c(n);
  for(size_t i=0;i<n;i++) a[i]=(char)(i*7);
  for (int thr : {1,2,4,8}) { set_host_copy_threads(thr);
   for (int nt : {0,1,2,3}) { set_host_copy_nt(nt);
    auto t0=std::chrono::steady_clock::now();
    for(int r=0;r<3;r++){ CopyJob j[2]={{b.data()+1,a.data()+3,n-5,false},{c.data(),a.data(),n,true}}; host_copy_parallel(j,2);}
    double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
    const int good = !memcmp(b.data()+1,a.data()+3,n-5)&&!memcmp(c.data(),a.data(),n); bad += !good;
    fprintf(stderr,"thr %d nt %d: %.1f GB/s ok=%d\n",thr,nt,2.0*n*3/dt/1e9,good); memset(b.data(),0,n); memset(c.data(),0,n);
   }}
  // concurrency: 4 callers
  std::vector<std::thread> th; for(int t=0;t<4;t++) th.emplace_back([&,t]{ std::vector<char> d(n/4); for(int r=0;r<20;r++){CopyJob j{d.data(),a.data()+t*(n/4),n/4,true}; host_copy_parallel(&j,1); if(memcmp(d.data(),a.data()+t*(n/4),n/4)) bad++;}});
  for(auto&t:th)t.join();
  host_copy_shutdown();
  // the shared chunk counter of xDct32BatchMultiGpu: 8 threads claiming concurrently cover every unit exactly once, ragged tail included
  for (size_t nUnits : {(size_t)1, (size_t)16384, (size_t)(16384 * 37 + 5), (size_t)1000003}) {
    ChunkClaimer cl(nUnits, nUnits < 16384 ? nUnits : 16384, true);
    std::vector<unsigned char> seen(nUnits, 0);
    std::atomic<size_t> chunks{0};
    std::vector<std::thread> cs;
    for (int t = 0; t < 8; t++) cs.emplace_back([&] { size_t u0, nu; while (cl.claim(&u0, &nu)) { chunks++; for (size_t u = u0; u < u0 + nu; u++) seen[u]++; } });
    for (auto& c : cs) c.join();
    for (size_t u = 0; u < nUnits; u++) if (seen[u] != 1) { bad++; break; }
    if (chunks != (nUnits + cl.chunk - 1) / cl.chunk) bad++;
  }
  puts(bad ? "FAILED" : "ok"); return bad != 0;
}
