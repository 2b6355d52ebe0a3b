/* xraylib.h — DECLARATIONS ONLY stand-in for xraylib 4.x's public header. TEST INFRASTRUCTURE.
 *
 * xraylib (T. Schoonjans et al., BSD licence) is the element data library DXMClib is built on; it is neither vendored by the
 * reference nor installed in this image. dxmclib_b200/host/matdb.cpp is written against its C API and selects it with
 * -DDXMCB200_USE_XRAYLIB. This header declares the subset of that API matdb.cpp calls, with xraylib 4's signatures (trailing
 * xrl_error**), so that tests/test_xraylib_path_compiles.py can type-check the xraylib code path on every run and it cannot
 * rot while no real xraylib is around. No definitions: nothing can link against it. */
#ifndef XRAYLIB_H_STUB
#define XRAYLIB_H_STUB

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { XRL_ERROR_MEMORY, XRL_ERROR_INVALID_ARGUMENT, XRL_ERROR_IO, XRL_ERROR_TYPE, XRL_ERROR_UNSUPPORTED, XRL_ERROR_RUNTIME } xrl_error_code;
typedef struct {
    xrl_error_code code;
    char* message;
} xrl_error;

#define K_SHELL 0
#define L1_SHELL 1
#define L2_SHELL 2
#define L3_SHELL 3
#define M1_SHELL 4
#define M2_SHELL 5
#define M3_SHELL 6
#define M4_SHELL 7
#define M5_SHELL 8
#define N1_SHELL 9
#define N2_SHELL 10
#define N3_SHELL 11
#define N4_SHELL 12
#define N5_SHELL 13
#define N6_SHELL 14
#define N7_SHELL 15
#define O1_SHELL 16
#define O2_SHELL 17
#define O3_SHELL 18
#define O4_SHELL 19
#define O5_SHELL 20
#define O6_SHELL 21
#define O7_SHELL 22
#define P1_SHELL 23
#define P2_SHELL 24
#define P3_SHELL 25

#define FL12_TRANS 1
#define FL13_TRANS 2
#define FLP13_TRANS 3
#define FL23_TRANS 4

struct compoundData {
    int nElements;
    double nAtomsAll;
    int* Elements;
    double* massFractions;
    double* nAtoms;
    double molarMass;
};
struct compoundDataNIST {
    char* name;
    int nElements;
    int* Elements;
    double* massFractions;
    double density;
};

double AtomicWeight(int Z, xrl_error** error);
double ElementDensity(int Z, xrl_error** error);
char* AtomicNumberToSymbol(int Z, xrl_error** error);
int SymbolToAtomicNumber(const char* symbol, xrl_error** error);
void xrlFree(void*);

double CS_Total(int Z, double E, xrl_error** error);
double CS_Photo(int Z, double E, xrl_error** error);
double CS_Rayl(int Z, double E, xrl_error** error);
double CS_Compt(int Z, double E, xrl_error** error);
double CS_Energy(int Z, double E, xrl_error** error);
double CS_Total_CP(const char compound[], double E, xrl_error** error);
double CS_Photo_CP(const char compound[], double E, xrl_error** error);
double CS_Rayl_CP(const char compound[], double E, xrl_error** error);
double CS_Compt_CP(const char compound[], double E, xrl_error** error);
double CS_Energy_CP(const char compound[], double E, xrl_error** error);

double FF_Rayl(int Z, double q, xrl_error** error);
double SF_Compt(int Z, double q, xrl_error** error);

double EdgeEnergy(int Z, int shell, xrl_error** error);
double ElectronConfig(int Z, int shell, xrl_error** error);
double ComptonProfile_Partial(int Z, int shell, double pz, xrl_error** error);
double FluorYield(int Z, int shell, xrl_error** error);
double CosKronTransProb(int Z, int trans, xrl_error** error);
double RadRate(int Z, int line, xrl_error** error);
double LineEnergy(int Z, int line, xrl_error** error);
double CSb_Photo_Partial(int Z, int shell, double E, xrl_error** error);

struct compoundData* CompoundParser(const char compoundString[], xrl_error** error);
void FreeCompoundData(struct compoundData* compoundData);
struct compoundDataNIST* GetCompoundDataNISTByName(const char compoundString[], xrl_error** error);
void FreeCompoundDataNIST(struct compoundDataNIST* compoundData);
char** GetCompoundDataNISTList(int* nCompounds, xrl_error** error);

#ifdef __cplusplus
}
#endif
#endif
