// xrd_sathelper.hpp -- header-only C++ host side above the C ABI (include/xrd.h).
//
// Mirrors the operator interface the reference's hot path is written against, so that the call
// sites of reference demodulator/src/demodulator.cpp compile unchanged against it:
//
//   SatHelper::Filters::RRC / lowPass                       demodulator.cpp:443-444
//   SatHelper::FirFilter(decimation, taps)::Work            demodulator.cpp:446,450,138,148
//   SatHelper::AGC(rate, reference, gain, maxGain)::Work    demodulator.cpp:447,143
//   SatHelper::CostasLoop(loopBandwidth, order)::Work       demodulator.cpp:448,152
//   SatHelper::ClockRecovery(omega, gainOmega, mu, gainMu, omegaRelativeLimit)::Work -> int
//                                                           demodulator.cpp:449,156
//   SatHelper::SatHelperException                           demodulator.cpp:499-501
//
// and, one level up, xrd::Demodulator: the whole processSamples() loop (demodulator.cpp:100-168)
// as one object with the reference's two seams -- addSamples(void*, int, int) is
// onSamplesAvailable (demodulator.cpp:54-74, the callback type of
// FrontendDevice::SetSamplesAvailableCallback, FrontendDevice.h:37) and the sink is anything with
// add(std::complex<float>*, int) (SymbolManager.h:37); an optional second sink with
// addSamples(const float*, int) receives what DiagManager::addSamples does (demodulator.cpp:161-163).
//
// Every Work() runs hand-written sm_100a kernels through libxrd.so; errors surface as
// SatHelperException exactly where the reference catches them.  There is no CPU path here.
#ifndef XRD_SATHELPER_HPP_
#define XRD_SATHELPER_HPP_

#include <complex>
#include <cstdint>
#include <exception>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "xrd.h"

namespace SatHelper {

class SatHelperException : public std::exception {
public:
    explicit SatHelperException(const std::string &r) : msg(r) {}
    const std::string &reason() const { return msg; }
    const char *what() const noexcept override { return msg.c_str(); }

private:
    std::string msg;
};

namespace FFTWindows {
enum WindowType { NONE = 0, HAMMING = 1, HANN = 2, BLACKMAN = 3, RECTANGULAR = 4, KAISER = 5, BLACKMAN_HARRIS = 6 };
}

class Filters {
public:
    // demodulator.cpp:443
    static std::vector<float> RRC(double gain, double sampleRate, double symbolRate, double alpha, int nTaps)
    {
        std::vector<float> t((size_t)(nTaps | 1));
        const int n = xrd_design_rrc(gain, sampleRate, symbolRate, alpha, nTaps, t.data(), (int)t.size());
        if (n < 0) throw SatHelperException("Filters::RRC: bad arguments");
        t.resize((size_t)n);
        return t;
    }
    // demodulator.cpp:444 -- only the Hamming window the reference asks for is implemented
    static std::vector<float> lowPass(double gain, double sampleRate, double cutFrequency, double transitionWidth,
                                      FFTWindows::WindowType window, double /*beta*/)
    {
        if (window != FFTWindows::HAMMING) throw SatHelperException("Filters::lowPass: only HAMMING is implemented");
        std::vector<float> t(1);
        int n = xrd_design_lowpass(gain, sampleRate, cutFrequency, transitionWidth, t.data(), 1);
        if (n < -1) {   // -n = taps needed
            t.resize((size_t)-n);
            n = xrd_design_lowpass(gain, sampleRate, cutFrequency, transitionWidth, t.data(), (int)t.size());
        }
        if (n < 0) throw SatHelperException("Filters::lowPass: bad arguments");
        t.resize((size_t)n);
        return t;
    }
};

namespace detail {
// one GPU stage operator (xrd_stage) with SatHelper's Work(in, out, length) call shape
class Stage {
public:
    Stage() = default;
    Stage(const Stage &) = delete;
    Stage &operator=(const Stage &) = delete;
    virtual ~Stage() { xrd_stage_destroy(h); }

protected:
    xrd_stage *h = nullptr;
    void created(int rc, const char *what)
    {
        if (rc != XRD_OK) throw SatHelperException(std::string(what) + ": " + xrd_stage_last_error(nullptr));
    }
    int work(std::complex<float> *in, std::complex<float> *out, int length)
    {
        const int rc = xrd_stage_work(h, reinterpret_cast<const float *>(in), reinterpret_cast<float *>(out), length);
        if (rc < 0) throw SatHelperException(xrd_stage_last_error(h));
        return rc;
    }
};
}  // namespace detail

class FirFilter : public detail::Stage {
public:
    FirFilter(unsigned int decimation, const std::vector<float> &taps, int device = 0)
    {
        created(xrd_fir_create(device, decimation, taps.data(), (int)taps.size(), &h), "FirFilter");
    }
    // length = number of OUTPUT samples; consumes length * decimation inputs (demodulator.cpp:137-138)
    void Work(std::complex<float> *input, std::complex<float> *output, int length) { work(input, output, length); }
};

class AGC : public detail::Stage {
public:
    AGC(float rate, float reference, float gain, float maxGain, int device = 0)
    {
        created(xrd_agc_create(device, rate, reference, gain, maxGain, &h), "AGC");
    }
    void Work(std::complex<float> *input, std::complex<float> *output, int length) { work(input, output, length); }
};

class CostasLoop : public detail::Stage {
public:
    CostasLoop(float loopBandwidth, int order, int device = 0)
    {
        created(xrd_costas_create(device, loopBandwidth, order, &h), "CostasLoop");
    }
    void Work(std::complex<float> *input, std::complex<float> *output, int length) { work(input, output, length); }
};

class ClockRecovery : public detail::Stage {
public:
    ClockRecovery(float omega, float gainOmega, float mu, float gainMu, float omegaRelativeLimit, int device = 0)
    {
        created(xrd_clock_recovery_create(device, omega, gainOmega, mu, gainMu, omegaRelativeLimit, &h), "ClockRecovery");
    }
    // returns the number of symbols written to output (demodulator.cpp:156)
    int Work(std::complex<float> *input, std::complex<float> *output, int length) { return work(input, output, length); }
};

}  // namespace SatHelper

namespace xrd {

// processSamples() + its globals (demodulator.cpp:31-52, 100-168) as one object.
class Demodulator {
public:
    // the sample callback type of FrontendDevice::SetSamplesAvailableCallback (FrontendDevice.h:37)
    typedef std::function<void(void *data, int length, int type)> SamplesCallback;

    // hrit: setHRITMode / setLRITMode presets (demodulator.cpp:177-197); edit config() before use
    explicit Demodulator(bool hrit, int device = 0, int channels = 1)
    {
        xrd_config_defaults(&cfg, hrit ? 1 : 0);
        cfg.device_ordinal = device;
        cfg.n_channels = channels;
    }
    explicit Demodulator(const xrd_config &c) : cfg(c) {}
    Demodulator(const Demodulator &) = delete;
    Demodulator &operator=(const Demodulator &) = delete;
    ~Demodulator() { xrd_destroy(h); }

    xrd_config &config() { return cfg; }

    // == onSamplesAvailable(void *fdata, int length, int type), demodulator.cpp:54-74.
    // Unknown types and FIFO overflow are reported on stderr by the reference and otherwise
    // ignored; here they come back as false with lastError() set.
    bool addSamples(void *data, int length, int type, int channel = 0)
    {
        open();
        return xrd_add_samples(h, channel, data, length, type) == XRD_OK;
    }
    // bind as the frontend's callback: device->SetSamplesAvailableCallback(demod.callback())
    SamplesCallback callback()
    {
        return [this](void *d, int n, int t) { this->addSamples(d, n, t); };
    }

    // == processSamples(), demodulator.cpp:100-168: runs the chain over everything queued (if at
    // least 32768 complex samples are, demodulator.cpp:113) and hands the symbols to
    // sink.add(std::complex<float>*, int) -- SymbolManager::add (SymbolManager.h:37).
    // Returns the number of complex samples consumed.
    template <class Sink> int64_t processSamples(Sink &sink, int64_t minSamples = 32768)
    {
        NoDiag none;
        return processSamples(sink, none, minSamples);
    }
    // The same with the reference's diagnostic tap (demodulator.cpp:161-163): after sink.add(ba, symbols),
    // diag.addSamples((float *)ba, symbols < 1024 ? symbols : 1024) -- DiagManager::addSamples (DiagManager.h:35)
    // sees the first min(symbols, 1024) FLOATS of the interleaved complex symbol buffer of every chunk.
    template <class Sink, class Diag> int64_t processSamples(Sink &sink, Diag &diag, int64_t minSamples = 32768)
    {
        open();
        struct Ctx { Sink *s; Diag *g; } ctx{&sink, &diag};
        const int64_t rc = xrd_process(
            h, minSamples,
            [](void *user, int /*channel*/, const float *sym, int n) {
                Ctx *c = static_cast<Ctx *>(user);
                c->s->add(reinterpret_cast<std::complex<float> *>(const_cast<float *>(sym)), n);
                c->g->addSamples(sym, n < 1024 ? n : 1024);
            },
            &ctx);
        if (rc < 0) throw SatHelper::SatHelperException(xrd_last_error(h));
        return rc;
    }

    // one-shot over a host buffer (state carried across calls); returns the symbol count
    int64_t demod(const std::complex<float> *iq, size_t n, std::complex<float> *symbols, size_t cap)
    {
        open();
        int64_t cnt[64] = {0};
        std::vector<int64_t> big;
        int64_t *c = cnt;
        if (cfg.n_channels > 64) {
            big.assign((size_t)cfg.n_channels, 0);
            c = big.data();
        }
        const int rc = xrd_demod_batch(h, iq, n, XRD_FLOATIQ, reinterpret_cast<float *>(symbols), cap, c);
        if (rc < 0) throw SatHelper::SatHelperException(xrd_last_error(h));
        return c[0];
    }

    // SymbolManager::process's byte rule (SymbolManager.cpp:43-46) for a block of symbols
    void softSymbols(const std::complex<float> *symbols, size_t n, int8_t *out)
    {
        open();
        if (xrd_soft_i8(h, reinterpret_cast<const float *>(symbols), n, out) < 0)
            throw SatHelper::SatHelperException(xrd_last_error(h));
    }

    xrd_loop_state state(int channel = 0)
    {
        open();
        xrd_loop_state st;
        if (xrd_get_state(h, channel, &st) < 0) throw SatHelper::SatHelperException(xrd_last_error(h));
        return st;
    }
    void setState(const xrd_loop_state &st, int channel = 0)
    {
        open();
        if (xrd_set_state(h, channel, &st) < 0) throw SatHelper::SatHelperException(xrd_last_error(h));
    }
    // checkpoint / resume of the whole demodulator (loop variables, filter histories, M&M tail, totals)
    std::vector<unsigned char> checkpoint()
    {
        open();
        std::vector<unsigned char> blob(xrd_checkpoint_size(h));
        if (xrd_checkpoint_save(h, blob.data(), blob.size()) < 0) throw SatHelper::SatHelperException(xrd_last_error(h));
        return blob;
    }
    void restore(const std::vector<unsigned char> &blob)
    {
        open();
        if (xrd_checkpoint_load(h, blob.data(), blob.size()) < 0) throw SatHelper::SatHelperException(xrd_last_error(h));
    }
    const char *lastError() const { return xrd_last_error(h); }
    xrd_demod *handle()
    {
        open();
        return h;
    }

    // Creates the device side now instead of on first use.  The reference wires addSamples to the frontend
    // thread and processSamples to the symbol-loop thread (demodulator.cpp:434,475); the first calls of the two
    // may race, so creation is serialised (std::call_once) whichever thread gets there first.
    void open()
    {
        std::call_once(once, [this]() {
            if (xrd_create(&cfg, &h) != XRD_OK) createError = xrd_last_error(nullptr);
        });
        if (!h) throw SatHelper::SatHelperException(createError.empty() ? "xrd_create failed" : createError);
    }

private:
    struct NoDiag {
        void addSamples(const float *, int) {}
    };
    xrd_config cfg;
    xrd_demod *h = nullptr;
    std::once_flag once;
    std::string createError;
};

}  // namespace xrd
#endif  // XRD_SATHELPER_HPP_
