// tests/emu/emu_probe.h -- included once by every harness: can this machine run a 1024-thread block as host threads?
#pragma once
#include <atomic>
#include <system_error>
#include <thread>
#include <vector>
extern "C" int emu_probe_threads(int n) {
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    bool ok = true;
    try {
        for (int i = 0; i < n; ++i) th.emplace_back([&] { while (!go.load()) std::this_thread::yield(); });
    } catch (const std::system_error&) { ok = false; }
    go.store(true);
    for (auto& t : th) t.join();
    return ok ? 1 : 0;
}
