// shim_test.cpp -- test driver for include/xrd_sathelper.hpp, built with plain g++ against
// libxrd.so (tests/test_shim.py calls it through ctypes).
//
// run_operator_chain() is written the way the reference's main() and processSamples() use the
// five operators (demodulator.cpp:436-450 construction, :132-159 ping-pong Work calls) so the test
// proves those call shapes compile and run against the B200 path; run_demodulator() drives the same
// stream through xrd::Demodulator's two seams (sample callback in, SymbolManager-style sink out).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <complex>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "xrd_sathelper.hpp"

using namespace SatHelper;

namespace {

std::string g_error;

// constants of demodulator/src/Parameters.h:27-37
const int RRC_TAPS = 63;
const int LOOP_ORDER = 2;
const float CLOCK_ALPHA = 0.0037f;
const float CLOCK_MU = 0.5f;
const float CLOCK_OMEGA_LIMIT = 0.005f;
const float CLOCK_GAIN_OMEGA = (CLOCK_ALPHA * CLOCK_ALPHA) / 4.0f;
const float AGC_RATE = 0.01f, AGC_REFERENCE = 0.5f, AGC_GAIN = 1.f, AGC_MAX_GAIN = 4000.f;

// stands in for OpenSatelliteProject::SymbolManager (SymbolManager.h:37): collects what add() is given
struct CollectingSink {
    std::vector<std::complex<float>> symbols;
    void add(std::complex<float> *data, int length) { symbols.insert(symbols.end(), data, data + length); }
};

// stands in for DiagManager (DiagManager.h:35, DiagManager.cpp:60-64): collects what addSamples() is given
struct CollectingDiag {
    std::vector<float> floats;
    void addSamples(const float *data, int length) { floats.insert(floats.end(), data, data + length); }
};

inline void swapBuffers(std::complex<float> **a, std::complex<float> **b) { std::swap(*a, *b); }

}  // namespace

extern "C" {

const char *shim_last_error(void) { return g_error.c_str(); }

// Five operators, constructed and called as the reference does; input is fed in `chunk`-sample
// calls (a multiple of the decimation).  Returns the number of symbols, or -1.
long long shim_run_operator_chain(const float *iq, long long n, unsigned sampleRate, unsigned symbolRate, float rrcAlpha,
                                  unsigned baseDecimation, int chunk, float *symOut, long long cap)
{
    try {
        float circuitSampleRate = sampleRate / ((float)baseDecimation);
        float sps = circuitSampleRate / ((float)symbolRate);
        float pllAlpha = CLOCK_ALPHA;   // demodulator.cpp:220

        std::vector<float> rrcTaps = Filters::RRC(1, circuitSampleRate, symbolRate, rrcAlpha, RRC_TAPS);
        std::vector<float> decimatorTaps =
            Filters::lowPass(1, sampleRate, circuitSampleRate / 2, 100e3, FFTWindows::WindowType::HAMMING, 6.76);

        FirFilter decimator(baseDecimation, decimatorTaps);
        AGC agc(AGC_RATE, AGC_REFERENCE, AGC_GAIN, AGC_MAX_GAIN);
        CostasLoop costasLoop(pllAlpha, LOOP_ORDER);
        ClockRecovery clockRecovery(sps, CLOCK_GAIN_OMEGA, CLOCK_MU, CLOCK_ALPHA, CLOCK_OMEGA_LIMIT);
        FirFilter rrcFilter(1, rrcTaps);
        CollectingSink symbolManager;

        std::vector<std::complex<float>> buffer0((size_t)chunk), buffer1((size_t)chunk);
        for (long long pos = 0; pos < n; pos += chunk) {
            int length = (int)std::min<long long>(chunk, n - pos);
            memcpy(buffer0.data(), iq + 2 * pos, sizeof(float) * 2 * (size_t)length);
            std::complex<float> *ba = buffer0.data(), *bb = buffer1.data();
            if (baseDecimation > 1) {
                length /= baseDecimation;
                decimator.Work(ba, bb, length);
                swapBuffers(&ba, &bb);
            }
            agc.Work(ba, bb, length);
            swapBuffers(&ba, &bb);
            rrcFilter.Work(ba, bb, length);
            swapBuffers(&ba, &bb);
            costasLoop.Work(ba, bb, length);
            swapBuffers(&ba, &bb);
            int symbols = clockRecovery.Work(ba, bb, length);
            swapBuffers(&ba, &bb);
            symbolManager.add(ba, symbols);
        }
        const long long ns = (long long)symbolManager.symbols.size();
        if (ns > cap) {
            g_error = "symbol capacity too small";
            return -1;
        }
        memcpy(symOut, symbolManager.symbols.data(), sizeof(float) * 2 * (size_t)ns);
        return ns;
    } catch (SatHelperException &e) {
        g_error = e.reason();
        return -1;
    }
}

// The same stream through xrd::Demodulator: a frontend-style callback delivers `block`-sample
// buffers of `type` (CFileFrontend delivers 65535, CFileFrontend.cpp:12,48), processSamples()
// drains the FIFO into the sink.  `raw` holds samples of `type` (FrontendDevice.h:11-13).
long long shim_run_demodulator(const void *raw, long long n, int type, int hrit, int block, float *symOut, long long cap)
{
    try {
        xrd::Demodulator demod(hrit != 0);
        CollectingSink symbolManager;
        xrd::Demodulator::SamplesCallback cb = demod.callback();   // what SetSamplesAvailableCallback receives
        const size_t bytes = (type == XRD_FLOATIQ) ? 8 : (type == XRD_S16IQ ? 4 : 2);
        for (long long pos = 0; pos < n; pos += block) {
            const int length = (int)std::min<long long>(block, n - pos);
            cb((void *)((const char *)raw + bytes * (size_t)pos), length, type);
            const bool last = pos + block >= n;
            demod.processSamples(symbolManager, last ? 1 : 32768);
        }
        const long long ns = (long long)symbolManager.symbols.size();
        if (ns > cap) {
            g_error = "symbol capacity too small";
            return -1;
        }
        memcpy(symOut, symbolManager.symbols.data(), sizeof(float) * 2 * (size_t)ns);
        return ns;
    } catch (SatHelperException &e) {
        g_error = e.reason();
        return -1;
    }
}

// The reference's thread wiring (demodulator.cpp:434,472-475): the frontend thread fires the sample callback, the
// symbol-loop thread polls processSamples(); the first calls of the two race to create the device side.  A
// DiagManager-style tap runs beside the sink (demodulator.cpp:161-163).  With threads == 0 the two alternate on
// the calling thread, one processSamples() per callback, which makes the taps deterministic.  Half way through,
// the demodulator is checkpointed and a NEW demodulator restored from the blob carries on (threads == 0 only).
long long shim_run_wired(const void *raw, long long n, int type, int hrit, int block, int threads, int checkpoint_at_block,
                         float *symOut, long long cap, float *diagOut, long long diagCap, long long *nDiag)
{
    try {
        std::unique_ptr<xrd::Demodulator> demod(new xrd::Demodulator(hrit != 0));
        CollectingSink symbolManager;
        CollectingDiag diagManager;
        const size_t bytes = (type == XRD_FLOATIQ) ? 8 : (type == XRD_S16IQ ? 4 : 2);
        if (threads) {
            std::atomic<bool> produced{false};
            std::string producerError;
            std::thread frontend([&]() {
                try {
                    xrd::Demodulator::SamplesCallback cb = demod->callback();
                    for (long long pos = 0; pos < n; pos += block) {
                        const int length = (int)std::min<long long>(block, n - pos);
                        // a real frontend drops on overflow (demodulator.cpp:104-106); the test waits instead
                        while (!demod->addSamples((void *)((const char *)raw + bytes * (size_t)pos), length, type))
                            std::this_thread::sleep_for(std::chrono::microseconds(200));
                    }
                } catch (SatHelperException &e) {
                    producerError = e.reason();
                }
                produced = true;
            });
            std::string loopError;
            std::thread symbolThread([&]() {   // symbolLoopFunc, demodulator.cpp:170-175
                try {
                    while (!produced) {
                        demod->processSamples(symbolManager, diagManager);
                        std::this_thread::sleep_for(std::chrono::microseconds(1));
                    }
                    while (demod->processSamples(symbolManager, diagManager, 1) > 0) {
                    }
                } catch (SatHelperException &e) {
                    loopError = e.reason();
                }
            });
            frontend.join();
            symbolThread.join();
            if (!producerError.empty() || !loopError.empty()) {
                g_error = producerError + loopError;
                return -1;
            }
        } else {
            int blk = 0;
            for (long long pos = 0; pos < n; pos += block, blk++) {
                if (blk == checkpoint_at_block && blk > 0) {
                    std::vector<unsigned char> blob = demod->checkpoint();
                    std::unique_ptr<xrd::Demodulator> next(new xrd::Demodulator(hrit != 0));
                    next->restore(blob);
                    demod.swap(next);   // the old demodulator is destroyed; the restored one carries on
                }
                const int length = (int)std::min<long long>(block, n - pos);
                demod->callback()((void *)((const char *)raw + bytes * (size_t)pos), length, type);
                demod->processSamples(symbolManager, diagManager, 1);
            }
        }
        const long long ns = (long long)symbolManager.symbols.size();
        const long long nd = (long long)diagManager.floats.size();
        if (ns > cap || nd > diagCap) {
            g_error = "output capacity too small";
            return -1;
        }
        memcpy(symOut, symbolManager.symbols.data(), sizeof(float) * 2 * (size_t)ns);
        memcpy(diagOut, diagManager.floats.data(), sizeof(float) * (size_t)nd);
        *nDiag = nd;
        return ns;
    } catch (SatHelperException &e) {
        g_error = e.reason();
        return -1;
    }
}

// error behaviour: constructing an operator the path does not implement throws SatHelperException
int shim_error_paths(void)
{
    int seen = 0;
    try {
        CostasLoop c(0.0037f, 4);   // QPSK order: not implemented
    } catch (SatHelperException &) {
        seen |= 1;
    }
    try {
        Filters::lowPass(1, 10e6, 1.25e6, 100e3, FFTWindows::WindowType::KAISER, 6.76);
    } catch (SatHelperException &) {
        seen |= 2;
    }
    return seen;
}

}  // extern "C"
