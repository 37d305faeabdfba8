// integration/segments_main.cpp -- `turing_b200_segments`: IDR-segment-parallel encoding on one device.
//
// The reference's `--segment N` makes every N-th picture an IDR and restarts the coding structure there
// (turing/InputQueue.cpp:229-233, :270-286, :311-324): the segments of a sequence share nothing.  The reference still
// encodes them one after the other; here up to S of them are encoded at the same time by S instances of the reference's
// own encode() (turing/encode.cpp:577) in one process.  All instances post their motion searches, PU costs, intra sweeps
// and transform blocks to the ONE submission queue of the process (integration/turing_hooks.cpp), so the batches the
// device sees are S times as deep and the round-trip latency of one instance is hidden behind the others.
//
// The output is the reference's own: segment 0's stream as written, every later segment's stream without the parameter
// sets and the active-parameter-sets SEI an independent encode repeats (NAL types 32, 33, 34, 39), its first NAL unit
// carrying the four-byte start code of an access unit's first NAL -- byte for byte what `turing encode --segment N` writes
// for the whole sequence (checked by tests/test_gpu_batched_encoder.py against oracle/_ref/turing_ref).
//
// usage: turing_b200_segments --parallel-segments S <turing encode options incl. --segment N --frames F -o OUT> input
// Ranks of a multi-GPU job take every R-th segment: --segment-rank r --segment-ranks R (each rank writes OUT.rank<r>.seg<k>
// and rank 0's caller concatenates; see bench.py).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "turing_hooks_state.h"
#include <chrono>

int encode(int argc, const char *const argv[]); // turing/encode.cpp:577

namespace {

std::vector<char> readFile(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// a later segment's stream without VPS / SPS / PPS / prefix SEI; the first NAL kept gets a four-byte start code
void appendSegment(std::ofstream &out, const std::vector<char> &b, bool first)
{
    if (first)
    {
        out.write(b.data(), (std::streamsize)b.size());
        return;
    }
    std::vector<size_t> starts; // position of each 00 00 01
    for (size_t i = 0; i + 2 < b.size(); ++i)
        if (b[i] == 0 && b[i + 1] == 0 && b[i + 2] == 1)
        {
            starts.push_back(i);
            i += 2;
        }
    bool wroteOne = false;
    for (size_t k = 0; k < starts.size(); ++k)
    {
        const size_t p = starts[k];
        size_t begin = p > 0 && b[p - 1] == 0 ? p - 1 : p;
        size_t end = k + 1 < starts.size() ? starts[k + 1] : b.size();
        if (k + 1 < starts.size() && b[end - 1] == 0) --end; // the next unit's zero_byte
        if (p + 3 >= b.size()) break;
        const int type = (b[p + 3] >> 1) & 0x3f;
        if (type == 32 || type == 33 || type == 34 || type == 39) continue;
        if (!wroteOne && begin == p) out.put(0);
        wroteOne = true;
        out.write(b.data() + begin, (std::streamsize)(end - begin));
    }
}

} // namespace

int main(int argc, const char *argv[])
{
    int parallel = 4, rank = 0, ranks = 1, width = 0, height = 0, bitDepth = 8, internalBitDepth = 0;
    long frames = -1, segment = -1, seek = 0, clipFrames = 0;
    std::string output;
    std::vector<std::string> pass; // options handed to every instance unchanged
    for (int i = 1; i < argc; ++i)
    {
        const std::string a = argv[i];
        auto value = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--parallel-segments") parallel = atoi(value().c_str());
        else if (a == "--segment-rank") rank = atoi(value().c_str());
        else if (a == "--segment-ranks") ranks = atoi(value().c_str());
        else if (a == "--frames") frames = atol(value().c_str());
        else if (a == "--seek") seek = atol(value().c_str());
        else if (a == "--segment") segment = atol(value().c_str());
        else if (a == "-o" || a == "--output-file") output = value();
        else if (a == "--clip-frames") clipFrames = atol(value().c_str()); // benchmarking: the input holds this many frames, segments wrap around
        else
        {
            pass.push_back(a);
            if ((a == "--input-res" || a == "--bit-depth" || a == "--internal-bit-depth") && i + 1 < argc)
            {
                const std::string v = argv[++i];
                pass.push_back(v);
                if (a == "--input-res") sscanf(v.c_str(), "%dx%d", &width, &height);
                else if (a == "--bit-depth") bitDepth = atoi(v.c_str());
                else internalBitDepth = atoi(v.c_str());
            }
        }
    }
    if (frames <= 0 || segment <= 0 || output.empty() || parallel < 1 || ranks < 1 || rank < 0 || rank >= ranks)
    {
        std::cerr << "usage: turing_b200_segments --parallel-segments S [--segment-rank r --segment-ranks R] --segment N --frames F -o OUT "
                     "<turing encode options> input.yuv\n";
        return 2;
    }
    // device pictures: every instance keeps its DPB and a source + reconstruction per picture in flight
    // (with the search hooks off only source pictures go up: the pictures an instance has in flight, HVB_HOOKS bits 0-2)
    const char *hooks = getenv("HVB_HOOKS");
    const bool references = !hooks || (atoi(hooks) & 7);
    setenv("HVB_POOL_PICTURES", std::to_string(std::min(900, (references ? 48 : 20) * parallel + 16)).c_str(), 0);
    // the device session before the clock starts (CUDA context, page-locked buffers, the picture pool: a one-off)
    if (width > 0 && height > 0 && hvbhooks::on())
    {
        const int depth = internalBitDepth ? internalBitDepth : bitDepth;
        hvbhooks::session(depth > 8 ? 2 : 1, depth, width, height);
    }
    const auto t0 = std::chrono::steady_clock::now();
    const long nSegments = (frames + segment - 1) / segment;
    std::vector<long> mine;
    for (long k = rank; k < nSegments; k += ranks) mine.push_back(k);
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto worker = [&] {
        for (;;)
        {
            const size_t at = next.fetch_add(1);
            if (at >= mine.size()) return;
            const long k = mine[at];
            const long first = k * segment, count = std::min(segment, frames - first);
            long from = seek + first;
            if (clipFrames > 0) from = (from / segment) % std::max(1L, clipFrames / segment) * segment;
            std::vector<std::string> args = {"turing encode", "--segment", std::to_string(segment), "--frames", std::to_string(count), "--seek",
                                             std::to_string(from), "-o", output + ".seg" + std::to_string(k)};
            args.insert(args.end(), pass.begin(), pass.end());
            std::vector<const char *> av;
            for (auto &s : args) av.push_back(s.c_str());
            if (encode((int)av.size(), av.data()) != 0) failed = 1;
        }
    };
    std::vector<std::thread> threads;
    for (int t = 0; t < std::min<long>(parallel, (long)mine.size()); ++t) threads.emplace_back(worker);
    for (auto &t : threads) t.join();
    if (failed) return 1;
    long encoded = 0;
    for (long k : mine) encoded += std::min(segment, frames - k * segment);
    fprintf(stderr, "segments wall: %.6f s for %ld frames in %zu segments (session set-up excluded)\n",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), encoded, mine.size());
    if (ranks == 1)
    {
        std::ofstream out(output, std::ios::binary);
        for (long k = 0; k < nSegments; ++k)
        {
            const std::string part = output + ".seg" + std::to_string(k);
            appendSegment(out, readFile(part), k == 0);
            remove(part.c_str());
        }
    }
    return 0;
}
