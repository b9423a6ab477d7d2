// cuFFT loaded with dlopen at first use (libeegldm.so has no link-time dependency on it); shared by spectral.cu (JukeboxLoss)
// and psd.cu (multitaper / Welch PSD of the sampling output tail).
#pragma once
#include <cufft.h>
#include <dlfcn.h>

#include <string>

namespace eegldm {

struct CufftApi {
    void* lib = nullptr;
    cufftResult (*PlanMany)(cufftHandle*, int, int*, int*, int, int, int*, int, int, cufftType, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecR2C)(cufftHandle, cufftReal*, cufftComplex*) = nullptr;
    cufftResult (*ExecC2R)(cufftHandle, cufftComplex*, cufftReal*) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) { err = "cannot dlopen libcufft.so.11"; return false; }
        PlanMany = (decltype(PlanMany))dlsym(lib, "cufftPlanMany");
        SetStream = (decltype(SetStream))dlsym(lib, "cufftSetStream");
        ExecR2C = (decltype(ExecR2C))dlsym(lib, "cufftExecR2C");
        ExecC2R = (decltype(ExecC2R))dlsym(lib, "cufftExecC2R");
        Destroy = (decltype(Destroy))dlsym(lib, "cufftDestroy");
        if (!PlanMany || !SetStream || !ExecR2C || !ExecC2R || !Destroy) { err = "libcufft is missing symbols"; return false; }
        return true;
    }
};

CufftApi& cufft_api();   // process-wide instance (spectral.cu); call load() under the caller's lock

}  // namespace eegldm
