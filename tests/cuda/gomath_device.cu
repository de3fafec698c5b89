// Evaluates the device restatements of the Go runtime functions (csrc/device/models.cuh: tsb_go_sin, tsb_go_max,
// tsb_go_min, tsb_go_pow, tsb_qdiv in its strict form) on inputs read from stdin and prints the results as hex
// doubles; tests/test_gomath_device.py compares them bit for bit with the oracle's host restatements.
// stdin: lines "<fn> <x> [<y>]" with fn in {sin, max, min, pow}; stdout: one "%a" per line.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../toy-spice_b200/csrc/device/models.cuh"

__global__ void eval(const int* fn, const double* x, const double* y, double* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (fn[i]) {
    case 0: out[i] = tsb_go_sin(x[i]); break;
    case 1: out[i] = tsb_go_max(x[i], y[i]); break;
    case 2: out[i] = tsb_go_min(x[i], y[i]); break;
    case 3: out[i] = tsb_go_pow(x[i], y[i]); break;
    default: out[i] = tsb_go_max_nn(x[i], y[i]); break;
    }
}

int main() {
    std::vector<int> fn; std::vector<double> x, y;
    char name[16]; double a, b;
    char line[256];
    while (fgets(line, sizeof line, stdin)) {
        b = 0.0;
        int k = sscanf(line, "%15s %la %la", name, &a, &b);
        if (k < 2) continue;
        fn.push_back(!strcmp(name, "sin") ? 0 : !strcmp(name, "max") ? 1 : !strcmp(name, "min") ? 2 : !strcmp(name, "pow") ? 3 : 4);
        x.push_back(a); y.push_back(b);
    }
    int n = (int)fn.size();
    int* dfn; double *dx, *dy, *dout;
    if (cudaMalloc(&dfn, n * sizeof(int)) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 2; }
    cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&dout, n * 8);
    cudaMemcpy(dfn, fn.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), n * 8, cudaMemcpyHostToDevice);
    eval<<<(n + 127) / 128, 128>>>(dfn, dx, dy, dout, n);
    std::vector<double> out(n);
    if (cudaMemcpy(out.data(), dout, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 3; }
    for (int i = 0; i < n; ++i) printf("%a\n", out[i]);
    return 0;
}
