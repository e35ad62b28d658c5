// structs.cuh -- internal definitions of the opaque C-ABI handles.
#pragma once
#include <memory>

#include "common.cuh"

// Sliced-ELL geometry shared by the matrix of an equation and the per-cell link
// tables of a mesh: slice s holds rows [32 s, 32 s + 32); entry k of row r lives
// at  sliceOff[s] + k*32 + (r & 31).  Entry 0 is the diagonal, entry k >= 1 is
// interior link k-1 of the cell (reference link order, I2).  Padded entries
// have col = row and val = 0.
struct SellPattern {
  int nRows = 0;      // owned rows
  int nCols = 0;      // owned + ghost columns (device numbering)
  int nSlices = 0;
  long long nSlots = 0;   // padded slot count
  long long nnz = 0;      // true entries
  std::vector<int> hSliceOff;  // nSlices+1 (in slots)
  std::vector<int> hRowLen;    // entries per row (incl. diagonal)
  std::vector<int> hCol;       // nSlots
  phb::DevBuf<int> sliceOff, rowLen, col;
};

struct phb_mesh {
  phb_ctx *ctx = nullptr;
  bool finalized = false;
  // ---- host, reference numbering (I1, I2)
  int nNodes = 0, nCells = 0, nFaces = 0;
  std::vector<double> nodeX, nodeY;
  std::vector<int> cptr, cind;
  std::vector<int> fN1, fN2, fL, fR, fPatch;
  std::vector<double> fCx, fCy, fSx, fSy;   // S_f oriented out of lCell
  std::vector<double> fG, fW, fQx, fQy;     // g_f, lCell weight, r/|r|^2
  std::vector<double> vol, cCx, cCy;
  std::vector<int> ilPtr, ilFace, ilCell, blPtr, blFace, dlPtr, dlCell;
  std::vector<std::string> patchNames;
  std::vector<uint64_t> hKey;
  std::vector<int> hVal;
  // canonical CSR pattern [P, nb in link order] over owned rows, reference ids
  std::vector<int> rowPtr, colInd, slotL, slotR, slotDiag;
  // ---- parallel (I3, I5), reference conventions
  int rank = 0, nProcs = 1, nLocal = 0, rowOffset = 0;
  std::vector<int> owner, globalId, localRow, globalRow;
  std::vector<int> bufPtr, bufCell, sendPtr, sendCell;
  // ---- device numbering: owned cells (IndexMap order) then ghosts by peer
  std::vector<int> cell2dev, dev2cell;
  int nDev = 0;  // nLocal + ghosts
  // ---- device SoA
  SellPattern sell;                 // scalar pattern over owned rows
  phb::DevBuf<int> dLinkFace;       // per slot: face id*2 + (cell is rCell), -1 for diag/pad
  phb::DevBuf<double> dVol;         // per dev cell
  phb::DevBuf<double> dFSx, dFSy, dFG, dFW, dFQx, dFQy;  // per face
  phb::DevBuf<int> dFL, dFR;        // per face, dev cell ids (-1 boundary)
  // boundary cells (owned cells with >= 1 boundary face)
  int nBCells = 0, nBFaces = 0;
  phb::DevBuf<int> dBcCell, dBcPtr, dBcFace;  // CSR over boundary cells -> faces
  phb::DevBuf<int> dBfFace, dBfCell, dBfPatch; // flat boundary face list
  phb::DevBuf<int> dIfFace;                   // interior face list
  int nIFaces = 0;
  // halo: send list (dev cell ids) grouped by peer; ghosts contiguous per peer
  std::vector<int> hSendDev, hSendCnt, hRecvCnt, hSendOff, hRecvOff;
  phb::DevBuf<int> dSendDev;
  phb::DevBuf<double> dSendBuf;
  // host<->device field permutation
  phb::DevBuf<int> dCell2Dev;
  phb::DevBuf<double> dStage;       // staging for phb_field_set/get when cell2dev is not the identity
  bool identityCells = false;       // cell2dev[i] == i and nDev == nCells: fields copy straight through
  phb::DevBuf<double> dCellC, dCo;  // cell centroids (device order) and Courant scratch, CICSAM only
  // peer-memory halo: where my values land in each peer's vectors (set by the launcher)
  std::vector<int> peerRecvOff, peerLd;
};

struct BcEntry {
  int type = PHB_NORMAL_GRADIENT;
  double vx = 0., vy = 0.;
};

struct phb_field {
  phb_mesh *m = nullptr;
  int nComp = 1;
  std::string name;
  std::vector<BcEntry> bc;              // per patch
  phb::DevBuf<int> dBfType;             // per boundary face (flat list order)
  phb::DevBuf<int> dFaceType;           // per face (interior faces: NORMAL_GRADIENT)
  bool bcDirty = true;
  unsigned bcVersion = 1;               // bumped by every boundary-condition change (matrix tags of the fused assembly)
  phb::DevBuf<double> cells, faces;     // [comp][nDev], [comp][nFaces]
  phb::DevBuf<double> cells0, faces0;   // old time level
  bool hasOld = false;
};

struct phb_eqn {
  phb_mesh *m = nullptr;
  int nComp = 1;
  phb::DevBuf<double> vals;   // SELL slots (coefficients shared by components)
  phb::DevBuf<double> rhs;    // [comp][nLocal]
  // vector equations with SYMMETRY patches: the 2 x 2 tensor the reference adds to the (cell, cell) block
  // (add(cell, cell, Tensor2D), UE/VectorFiniteVolumeEquation.cpp:48-66) on top of the shared coefficients:
  // [xx | xy | yx | yy][nLocal]; allocated on first use, hasTens = some entry may be non-zero
  phb::DevBuf<double> tens;
  bool hasTens = false;
};

struct phb_fracstep;
