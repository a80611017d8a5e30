// Rendezvous of DerpCLI's GPU worker threads (one per GPU) — kept in its own header so that the CPU test helper
// (IoSelfTest --mode=exchange) can exercise it without a GPU.
#pragma once
#include <condition_variable>
#include <mutex>
#include <vector>

// Rendezvous of the GPU worker threads when the destination cameras of a frame are dealt to several GPUs and
// the level handles mismatches: the stage reads EVERY camera's disparity (Derp.cpp:734-747), so the workers
// publish the device addresses of their planes, meet, copy the peers' planes GPU-to-GPU, meet again, update.
struct Exchange {
  explicit Exchange(int parties, int numCams) : parties(parties), planes(numCams, nullptr) {}
  void arriveAndWait() {
    std::unique_lock<std::mutex> lock(m);
    const int gen = generation;
    if (++waiting == parties) {
      waiting = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lock, [&] { return gen != generation; });
    }
  }
  const int parties;
  std::vector<const float*> planes;  // per rig camera: address of its disparity plane on the owning GPU
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0, generation = 0;
};
