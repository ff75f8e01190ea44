// Compile-only stand-ins for the ROOT classes the reference's spline sources mention.  TEST INFRASTRUCTURE
// (oracle/ref_host): every method is permissive and does nothing -- the code paths exercised through the harness
// (SMonolith built from TSpline3_red/TF1_red objects, Evaluate(); SampleHandlerBase::GetTestStatLLH) never reach
// ROOT.  A call that would need ROOT at run time aborts loudly (m3stub::no_root).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
typedef double Double_t; typedef int Int_t; typedef float Float_t; typedef bool Bool_t; typedef long long Long64_t;
typedef short Color_t; typedef unsigned int UInt_t; typedef short Short_t; typedef char Char_t; typedef const char Option_t;
namespace m3stub {
[[noreturn]] inline void no_root(const char* what) { std::fprintf(stderr, "ref_host stub: %s needs ROOT\n", what); std::abort(); }
struct Any {  // swallows any call chain
  template <class... A> Any(A&&...) {}
  template <class... A> Any operator()(A&&...) const { return Any(); }
  template <class T> operator T() const { no_root("value"); }
};
}  // namespace m3stub
#define M3STUB_CLASS(Name, Base)                                                     \
  class Name : public Base {                                                        \
   public:                                                                          \
    template <class... A> Name(A&&...) {}                                           \
    virtual ~Name() {}
#define M3STUB_END };
#define M3STUB_VOID(fn) template <class... A> void fn(A&&...) { m3stub::no_root(#fn); }
#define M3STUB_RET(T, fn) template <class... A> T fn(A&&...) const { m3stub::no_root(#fn); }
class TList; class TClass { public: M3STUB_RET(bool, InheritsFrom) };
class TObject { public: M3STUB_RET(TClass*, IsA) static TClass* Class() { return nullptr; } template <class... A> TObject(A&&...) {} virtual ~TObject() {}
  M3STUB_VOID(Write) M3STUB_VOID(SetName) M3STUB_VOID(SetTitle) M3STUB_VOID(Draw) M3STUB_VOID(Delete) M3STUB_VOID(SetDirectory)
  M3STUB_RET(const char*, GetName) M3STUB_RET(const char*, GetTitle) M3STUB_RET(const char*, ClassName) M3STUB_RET(TObject*, Clone) };
class TString : public std::string { public: using std::string::string; TString() {} TString(const std::string& s) : std::string(s) {}
  const char* Data() const { return c_str(); } int CompareTo(const char* o) const { return compare(o); } int CompareTo(const std::string& o) const { return compare(o); } bool Contains(const char* s) const { return find(s) != npos; }
  template <class... A> static TString Format(A&&...) { return TString(); } };
template <class... A> inline const char* Form(A&&...) { return ""; }
M3STUB_CLASS(TObjString, TObject) TString GetString() const { return TString(); } M3STUB_END
struct TArrayD { M3STUB_RET(const double*, GetArray) M3STUB_RET(double, GetAt) M3STUB_RET(int, GetSize) };
M3STUB_CLASS(TAxis, TObject) M3STUB_RET(const TArrayD*, GetXbins) M3STUB_VOID(SetBinLabel) M3STUB_RET(int, GetNbins) M3STUB_RET(double, GetBinLowEdge) M3STUB_RET(double, GetBinUpEdge) M3STUB_RET(int, FindBin) M3STUB_RET(int, FindFixBin) M3STUB_RET(double, GetBinCenter) M3STUB_RET(double, GetXmin) M3STUB_RET(double, GetXmax) M3STUB_VOID(SetTitle) M3STUB_END
M3STUB_CLASS(TGraph, TObject) M3STUB_RET(int, GetN) M3STUB_RET(int, GetPoint) M3STUB_VOID(SetPoint) M3STUB_VOID(Set) M3STUB_RET(double*, GetX) M3STUB_RET(double*, GetY) M3STUB_VOID(Fit) M3STUB_RET(double, Eval) M3STUB_END
M3STUB_CLASS(TSpline3, TObject) M3STUB_VOID(SetPointCoeff) M3STUB_VOID(SetNameTitle) M3STUB_RET(int, GetNp) M3STUB_VOID(GetKnot) M3STUB_VOID(GetCoeff) M3STUB_RET(double, Eval) M3STUB_RET(double, GetXmin) M3STUB_RET(double, GetXmax) M3STUB_END
M3STUB_CLASS(TSpline5, TObject) M3STUB_RET(int, GetNp) M3STUB_VOID(GetKnot) M3STUB_VOID(GetCoeff) M3STUB_RET(double, Eval) M3STUB_END
M3STUB_CLASS(TF1, TObject) M3STUB_RET(int, GetNpar) M3STUB_RET(double, GetParameter) M3STUB_VOID(SetParameter) M3STUB_RET(double, Eval) M3STUB_RET(TString, GetExpFormula) M3STUB_RET(double, GetXmin) M3STUB_RET(double, GetXmax) M3STUB_END
M3STUB_CLASS(TH1, TObject) M3STUB_VOID(SetFillColor) M3STUB_VOID(SetLineColor) M3STUB_VOID(SetLineWidth) M3STUB_VOID(SetLineStyle) M3STUB_VOID(Add) M3STUB_RET(TAxis*, GetXaxis) M3STUB_RET(TAxis*, GetYaxis) M3STUB_RET(TAxis*, GetZaxis) M3STUB_VOID(Fill) M3STUB_VOID(SetBinContent) M3STUB_RET(double, GetBinContent) M3STUB_RET(int, GetNbinsX) M3STUB_RET(double, Integral) M3STUB_VOID(Reset) M3STUB_VOID(Scale) M3STUB_END
M3STUB_CLASS(TH1D, TH1) M3STUB_END  M3STUB_CLASS(TH1F, TH1) M3STUB_END
M3STUB_CLASS(TH2, TH1) M3STUB_RET(int, GetNbinsY) M3STUB_END  M3STUB_CLASS(TH2D, TH2) M3STUB_END  M3STUB_CLASS(TH2F, TH2) M3STUB_END
M3STUB_CLASS(TH2Poly, TH2) M3STUB_RET(int, AddBin) M3STUB_RET(TList*, GetBins) M3STUB_RET(int, GetNumberOfBins) M3STUB_END  M3STUB_CLASS(TH3, TH2) M3STUB_END  M3STUB_CLASS(TH3D, TH3) M3STUB_END  M3STUB_CLASS(TH3F, TH3) M3STUB_END
M3STUB_CLASS(TBranch, TObject) M3STUB_VOID(SetAddress) M3STUB_VOID(GetEntry) M3STUB_END
M3STUB_CLASS(TTree, TObject) M3STUB_RET(TBranch*, Branch) M3STUB_VOID(SetBranchAddress) M3STUB_VOID(SetBranchStatus) M3STUB_VOID(GetEntry) M3STUB_RET(Long64_t, GetEntries) M3STUB_VOID(Fill) M3STUB_RET(TBranch*, GetBranch) M3STUB_VOID(ResetBranchAddresses) M3STUB_END
M3STUB_CLASS(TChain, TTree) M3STUB_VOID(Add) M3STUB_END
M3STUB_CLASS(TList, TObject) M3STUB_RET(int, GetSize) M3STUB_RET(TObject*, At) TObject** begin() const { m3stub::no_root("TList"); } TObject** end() const { m3stub::no_root("TList"); } M3STUB_END
M3STUB_CLASS(TKey, TObject) template <class T> T* ReadObject() { m3stub::no_root("TKey::ReadObject"); } M3STUB_RET(TObject*, ReadObj) M3STUB_RET(const char*, GetClassName) M3STUB_END
M3STUB_CLASS(TDirectory, TObject) M3STUB_VOID(cd) M3STUB_RET(TDirectory*, mkdir) M3STUB_RET(TObject*, Get) M3STUB_RET(TList*, GetListOfKeys) M3STUB_VOID(Close) M3STUB_RET(bool, IsOpen) M3STUB_RET(TDirectory*, GetDirectory) M3STUB_VOID(ls)
  template <class T> T* Get(const char*) { m3stub::no_root("Get<T>"); } M3STUB_END
M3STUB_CLASS(TDirectoryFile, TDirectory) M3STUB_END
M3STUB_CLASS(TFile, TDirectory) template <class... A> static TFile* Open(A&&...) { m3stub::no_root("TFile::Open"); } M3STUB_RET(bool, IsZombie) M3STUB_END
M3STUB_CLASS(TIterator, TObject) M3STUB_RET(TObject*, Next) M3STUB_END
class TIter { public: template <class... A> TIter(A&&...) {} TObject* operator()() { m3stub::no_root("TIter"); } TObject* Next() { m3stub::no_root("TIter"); } };
M3STUB_CLASS(TStopwatch, TObject) M3STUB_VOID(Start) M3STUB_VOID(Stop) M3STUB_RET(double, RealTime) M3STUB_RET(double, CpuTime) M3STUB_VOID(Reset) M3STUB_END
M3STUB_CLASS(TMacro, TObject) M3STUB_RET(TList*, GetListOfLines) M3STUB_VOID(AddLine) M3STUB_END
M3STUB_CLASS(TCanvas, TObject) M3STUB_VOID(Print) M3STUB_VOID(cd) M3STUB_VOID(SetGrid) M3STUB_END
M3STUB_CLASS(TPad, TObject) M3STUB_END
M3STUB_CLASS(TStyle, TObject) M3STUB_VOID(SetOptStat) M3STUB_END
M3STUB_CLASS(TRandom3, TObject) M3STUB_RET(double, Gaus) M3STUB_RET(double, Rndm) M3STUB_RET(double, Uniform) M3STUB_RET(int, Poisson) M3STUB_RET(double, PoissonD) M3STUB_VOID(SetSeed) M3STUB_END
M3STUB_CLASS(TROOT, TObject) M3STUB_RET(TClass*, GetClass) M3STUB_END
M3STUB_CLASS(TLegend, TObject) M3STUB_VOID(AddEntry) M3STUB_VOID(SetBorderSize) M3STUB_VOID(SetFillStyle) M3STUB_END
M3STUB_CLASS(THStack, TObject) M3STUB_VOID(Add) M3STUB_RET(TList*, GetHists) M3STUB_END
class TArrow; class TBox; class TCandle; class TColor; class TDecompChol; class TDecompSVD; class TEllipse; class TFitResult;
class TFitResultPtr; class TGraphAsymmErrors; class TGraphErrors; class TKDE; class TLatex; class TLine;
class TLorentzVector; class TMarker; class TMatrixDEigen; class TMatrixDSym; class TMatrixDSymEigen; class TPaveText; class TProfile;
class TText; class TVector3; class TVectorD; class TMatrixD;
#define ClassDef(name, version) static_assert(true, "")
namespace M3 { inline void AddPath(std::string&) {} }
extern TStyle* gStyle; extern TDirectory* gDirectory; extern TROOT* gROOT;
namespace TMath { inline double Log(double x); inline double Sqrt(double x); }
