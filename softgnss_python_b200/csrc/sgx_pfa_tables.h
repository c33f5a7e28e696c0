// Generated (see tools/gen_fft_tables.py, radix-31 section): constants of the grouped radix-31 butterfly of sgx_pfa.cu.
// 3 is a primitive root mod 31; input pair n is (x[J[n]], x[31-J[n]]) with J[n] = |3^n mod 31| folded into 1..15 and
// sign SG[n]; C31R[e] = cos(2 pi 3^e / 31), S31R[e] = sin(2 pi 3^e / 31) (e >= 15: minus the value 15 earlier);
// output pair m = 5 q + r goes to rows KHI31[m] and 31 - KHI31[m].
#pragma once
namespace sgx { namespace pfa {
#define SGX_R31_J {1, 3, 9, 4, 12, 5, 15, 14, 11, 2, 6, 13, 8, 7, 10}
#define SGX_R31_KHI {1, 3, 9, 27, 19, 26, 16, 17, 20, 29, 25, 13, 8, 24, 10}
#define SGX_R31_SG {1, 1, 1, -1, -1, -1, -1, -1, -1, -1, -1, 1, 1, -1, 1}
SGX_TABLE float C31R[15] = {9.795299172e-01f, 8.207634687e-01f, -2.506525218e-01f, 6.889669299e-01f, -7.587581277e-01f, 5.289639831e-01f, -9.948693514e-01f, -9.541392326e-01f, -6.121059656e-01f, 9.189578295e-01f, 3.473052382e-01f, -8.743466139e-01f, -5.064916983e-02f, 1.514277756e-01f, -4.403941631e-01f};
SGX_TABLE float S31R[30] = {2.012985200e-01f, 5.712682009e-01f, 9.680771232e-01f, -7.247927785e-01f, -6.513724923e-01f, -8.486442566e-01f, -1.011683196e-01f, -2.993631363e-01f, -7.907757163e-01f, -3.943558633e-01f, -9.377521276e-01f, 4.853019714e-01f, 9.987165332e-01f, -9.884683490e-01f, 8.978045583e-01f, -2.012985200e-01f, -5.712682009e-01f, -9.680771232e-01f, 7.247927785e-01f, 6.513724923e-01f, 8.486442566e-01f, 1.011683196e-01f, 2.993631363e-01f, 7.907757163e-01f, 3.943558633e-01f, 9.377521276e-01f, -4.853019714e-01f, -9.987165332e-01f, 9.884683490e-01f, -8.978045583e-01f};
SGX_TABLE int KHI31[15] = {1, 3, 9, 27, 19, 26, 16, 17, 20, 29, 25, 13, 8, 24, 10};
}}  // namespace sgx::pfa
