"""cask_b200/csrc/hostcopy.hpp (host threads that stage pageable vectors through pinned rings): the copy pool on the CPU —
every byte lands once, odd sizes and unaligned ends included, with two threads driving the pool at the same time as the
upload side and the drain side of cask_b200_spmv do."""
import os
import subprocess

from conftest import ROOT

SRC = r'''
#include "hostcopy.hpp"
#include <cstdio>
#include <cstdlib>
using namespace caskb200;
static int check(HostCopyPool& pool, size_t bytes, unsigned seed) {
  std::vector<unsigned char> src(bytes + 64), dst(bytes + 128, 0xAB);
  for (size_t i = 0; i < src.size(); i++) src[i] = (unsigned char)((i * 2654435761u + seed) >> 13);
  pool.copy(dst.data() + 7, src.data() + 3, bytes);   // unaligned on both sides
  for (size_t i = 0; i < bytes; i++) if (dst[7 + i] != src[3 + i]) return 1;
  for (size_t i = 0; i < 7; i++) if (dst[i] != 0xAB) return 2;
  for (size_t i = 7 + bytes; i < dst.size(); i++) if (dst[i] != 0xAB) return 3;   // nothing written past the end
  return 0;
}
int main() {
  for (int threads : {1, 3, 8}) {
    HostCopyPool pool(threads);
    if (pool.threads() != threads) return 10;
    const size_t sizes[] = {0, 1, 63, 4096, (256u << 10) - 1, (256u << 10), (1u << 20) + 17, (4u << 20), (9u << 20) + 5};
    std::atomic<int> bad{0};
    auto body = [&](unsigned seed) { for (int rep = 0; rep < 3; rep++) for (size_t s : sizes) { int r = check(pool, s, seed + rep); if (r) bad = r; } };
    std::thread a(body, 1u), b(body, 1000u);   // upload side and drain side share the pool
    a.join(); b.join();
    if (bad) { std::printf("FAILED %d with %d threads\n", bad.load(), threads); return 1; }
  }
  std::printf("OK\n");
  return 0;
}
'''


def test_host_copy_pool(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = str(tmp_path / "t")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "cask_b200", "csrc"),
                           "-I/usr/local/cuda/include", str(src), "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout
