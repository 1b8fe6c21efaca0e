{-# LANGUAGE ForeignFunctionInterface, RecordWildCards, ScopedTypeVariables #-}

-- | Drop-in GPU path for the numeric core of "QAP" (sdiehl/arithmetic-circuits, src/QAP.hs).
--
-- This module is the reference-side binding of the C ABI in @include/acg.h@ (libacg.so, CUDA sm_100a).
-- It re-exports module "QAP" and REPLACES, with the reference's own names and types monomorphised to
-- BN254 'Fr' (the FFI cannot be class-polymorphic; a second copy over BLS12-381 'Fr' differs only in
-- 'fieldId'), every function whose numeric work the library does:
--
-- > verifyAssignment      :: QAP Fr -> QapSet Fr -> Bool                              -- src/QAP.hs:276-282
-- > verificationWitness   :: QAP Fr -> QapSet Fr -> Maybe (VPoly Fr)                  -- :292-298
-- > verificationWitnessZk :: Fr -> Fr -> Fr -> QAP Fr -> QapSet Fr -> Maybe (VPoly Fr) -- :300-327  acg_qap_verify
-- > createPolynomials     :: GenQAP (Map Fr) Fr -> QAP Fr                             -- :486-508  acg_lagrange
-- > createPolynomialsFFT  :: (Int -> Fr) -> GenQAP (Map Fr) Fr -> QAP Fr              -- :512-525  acg_interpolate_columns
-- > arithCircuitToQAP     :: [[Fr]] -> ArithCircuit Fr -> QAP Fr                      -- :542-549
-- > arithCircuitToQAPFFT  :: (Int -> Fr) -> [[Fr]] -> ArithCircuit Fr -> QAP Fr       -- :552-561
-- > gateToQAP             :: (Int -> Fr) -> [Fr] -> Gate Wire Fr -> QAP Fr            -- :355-362
--
-- so that the reference's own test-suite type-checks against it unchanged (test/Test/QAP.hs:68-103 pass the
-- @QAP Fr@ built by 'arithCircuitToQAP' / 'gateToQAP' to 'verifyAssignment').  The R1CS-form fast path, which never
-- materialises the per-wire polynomials, is exported under NEW names:
--
-- > verifyAssignmentR1CS      :: GenQAP (Map Fr) Fr -> QapSet Fr -> Bool              -- acg_r1cs_check_host
-- > verificationWitnessZkR1CS :: Fr -> Fr -> Fr -> GenQAP (Map Fr) Fr -> QapSet Fr -> Maybe (VPoly Fr)   -- acg_qap_witness
--
-- Prerequisite in the reference: the functions that take a 'GenQAP' need its constructor and fields, which
-- module QAP does not export (src/QAP.hs:12-39 lists only 'arithCircuitToGenQAP').  hs/patches/QAP-export-GenQAP.patch
-- adds @GenQAP(..)@ and @createMapGenQap@ to that export list -- one line each, no code change.  The @QAP@-typed
-- functions need no patch.
--
-- NOTE: the build image of this repository has no GHC (SURVEY.md F3), so this file is shipped as source
-- and has not been compiled here; the same ABI is exercised from Python (ctypes) and C++ by the test
-- suite.  Marshalling follows include/acg.h: a field element is 4 little-endian Word64 limbs of @fromP@.
module QAP.GPU
  ( withAcg
  , verifyAssignment
  , verificationWitness
  , verificationWitnessZk
  , createPolynomials
  , createPolynomialsFFT
  , arithCircuitToQAP
  , arithCircuitToQAPFFT
  , gateToQAP
  , verifyAssignmentR1CS
  , verificationWitnessR1CS
  , verificationWitnessZkR1CS
  , module QAP
  ) where

import Protolude hiding (quot, quotRem)

import           Data.Bits                    (shiftL, shiftR, (.&.))
import qualified Data.Map                     as Map
import           Data.Pairing.BN254           (Fr, getRootOfUnity)
import           Data.Field.Galois            (fromP)
import           Data.Poly                    (VPoly, toPoly, unPoly)
import qualified Data.Vector                  as V
import qualified Data.Vector.Storable         as VS
import qualified Data.Vector.Storable.Mutable as VSM
import           Foreign.C.Types              (CInt(..), CUInt(..))
import           Foreign.Marshal.Alloc        (alloca, allocaBytes)
import           Foreign.Marshal.Array        (peekArray)
import           Foreign.Ptr                  (Ptr, nullPtr)
import           Foreign.Storable             (peek, pokeByteOff)
import           System.IO.Unsafe             (unsafePerformIO)

import           Circuit.Arithmetic           (ArithCircuit, Gate, Wire)
import           QAP hiding ( verifyAssignment, verificationWitness, verificationWitnessZk, createPolynomials
                            , createPolynomialsFFT, arithCircuitToQAP, arithCircuitToQAPFFT, gateToQAP )
import qualified QAP                          as Ref

-- opaque handles
data AcgCtx
data AcgR1cs
data AcgVec
data AcgQap

fieldId :: CInt
fieldId = 0 -- ACG_FIELD_BN254_FR

-- Calls are `safe`: they last far longer than a microsecond and block on the GPU.
foreign import ccall safe "acg_ctx_create"      c_ctx_create      :: CInt -> CInt -> Ptr (Ptr AcgCtx) -> IO CInt
foreign import ccall safe "acg_ctx_destroy"     c_ctx_destroy     :: Ptr AcgCtx -> IO ()
foreign import ccall safe "acg_r1cs_check_host" c_r1cs_check_host
  :: Ptr AcgCtx -> CUInt -> CUInt -> Ptr () -> Ptr () -> Ptr () -> Ptr Word64 -> Ptr Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "acg_r1cs_upload"     c_r1cs_upload
  :: Ptr AcgCtx -> CUInt -> CUInt -> Ptr () -> Ptr () -> Ptr () -> CUInt -> CUInt -> Ptr (Ptr AcgR1cs) -> IO CInt
foreign import ccall safe "acg_r1cs_free"       c_r1cs_free       :: Ptr AcgR1cs -> IO ()
foreign import ccall safe "acg_witness_upload"  c_witness_upload  :: Ptr AcgCtx -> Ptr Word64 -> CUInt -> Ptr (Ptr AcgVec) -> IO CInt
foreign import ccall safe "acg_vec_free"        c_vec_free        :: Ptr AcgVec -> IO ()
foreign import ccall safe "acg_qap_witness"     c_qap_witness
  :: Ptr AcgCtx -> Ptr AcgR1cs -> Ptr AcgVec -> Ptr Word64 -> Ptr Word64 -> Ptr Word64 -> Ptr Word64 -> Ptr Word64 -> Ptr CInt -> IO CInt
foreign import ccall safe "acg_interpolate_columns" c_interpolate_columns
  :: Ptr AcgCtx -> Ptr Word64 -> CUInt -> CUInt -> IO CInt
foreign import ccall safe "acg_lagrange"        c_lagrange
  :: Ptr AcgCtx -> Ptr Word64 -> Ptr Word64 -> CUInt -> CUInt -> Ptr Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "acg_fft_target"      c_fft_target      :: Ptr AcgCtx -> CUInt -> Ptr Word64 -> IO CInt
-- verificationWitnessZk on a QAP value, whole on the device (scale-and-sum, NTT product, long division by the target)
foreign import ccall safe "acg_qap_upload"      c_qap_upload
  :: Ptr AcgCtx -> Ptr Word64 -> Ptr Word64 -> Ptr Word64 -> CUInt -> CUInt -> Ptr Word64 -> CUInt -> Ptr (Ptr AcgQap) -> IO CInt
foreign import ccall safe "acg_qap_free"        c_qap_free        :: Ptr AcgQap -> IO ()
foreign import ccall safe "acg_qap_quotient_len" c_qap_quotient_len :: Ptr AcgQap -> IO CUInt
foreign import ccall safe "acg_qap_verify"      c_qap_verify
  :: Ptr AcgCtx -> Ptr AcgQap -> Ptr Word64 -> Ptr Word64 -> Ptr Word64 -> CUInt -> Ptr CUInt -> Ptr CInt -> IO CInt

-- | One context per device; the reference is single-threaded, so a process-wide context is enough.
withAcg :: Int -> (Ptr AcgCtx -> IO a) -> IO a
withAcg device k = alloca $ \pp -> do
  rc <- c_ctx_create fieldId (fromIntegral device) pp
  when (rc /= 0) $ panic ("acg_ctx_create failed: " <> show rc) -- no CPU fallback by design
  ctx <- peek pp
  k ctx `finally` c_ctx_destroy ctx

{-# NOINLINE globalCtx #-}
globalCtx :: Ptr AcgCtx
globalCtx = unsafePerformIO $ alloca $ \pp -> do
  rc <- c_ctx_create fieldId 0 pp
  when (rc /= 0) $ panic ("acg_ctx_create failed: " <> show rc)
  peek pp

ok :: Text -> CInt -> IO ()
ok what rc = when (rc /= 0) $ panic (what <> " failed: " <> show rc)

-- | @fromP@ as 4 little-endian limbs.
limbs :: Fr -> [Word64]
limbs x = [ fromIntegral ((n `shiftR` (64 * i)) .&. 0xFFFFFFFFFFFFFFFF) | i <- [0 .. 3] ]
  where n = toInteger (fromP x)

unlimbs :: [Word64] -> Fr
unlimbs ws = fromInteger $ sum [ toInteger w `shiftL` (64 * i) | (w, i) <- zip ws [0 ..] ]

chunk :: Int -> [a] -> [[a]]
chunk _ [] = []
chunk k xs = let (h, t) = splitAt k xs in h : chunk k t

frVector :: [Fr] -> VS.Vector Word64
frVector = VS.fromList . concatMap limbs

-- | Witness index layout of qapSetToMap (src/QAP.hs:605-620): block sizes = max key + 1 over everything paired.
data Layout = Layout { nIn, nMid, nOut :: Int }

layoutOf :: [QapSet ()] -> Layout
layoutOf qs = Layout (mx qapSetInput) (mx qapSetIntermediate) (mx qapSetOutput)
  where mx f = maximum (0 : [ k + 1 | q <- qs, k <- Map.keys (f q) ])

nCols :: Layout -> Int
nCols Layout{..} = 1 + nIn + nMid + nOut

-- | A QapSet as a dense list in qapSetToMap order; wires it lacks get the default (0 / the zero polynomial: exactly
-- what combineWithDefaults does on either side, src/QAP.hs:163-181, 314).
dense :: a -> Layout -> QapSet a -> [a]
dense def Layout{..} QapSet{..} =
  qapSetConstant : [ Map.findWithDefault def k qapSetInput | k <- [0 .. nIn - 1] ]
                ++ [ Map.findWithDefault def k qapSetIntermediate | k <- [0 .. nMid - 1] ]
                ++ [ Map.findWithDefault def k qapSetOutput | k <- [0 .. nOut - 1] ]

-- | Give the wires of a template QapSet, in Foldable order (constant, inputs, intermediates, outputs -- the order
-- @toList@ flattened them in), the next values of a list; returns the rest of the list.
refillSet :: QapSet b -> [a] -> (QapSet a, [a])
refillSet _ [] = panic "QAP.GPU.refillSet: ran out of polynomials"
refillSet (QapSet _ inp mid outp) (c : xs) = (QapSet c (fill inp is) (fill mid ms) (fill outp os), rest)
  where
    (is, r1)   = splitAt (Map.size inp) xs
    (ms, r2)   = splitAt (Map.size mid) r1
    (os, rest) = splitAt (Map.size outp) r2
    fill m vs  = Map.fromDistinctAscList (zip (Map.keys m) vs)

coeffs :: VPoly Fr -> [Fr]
coeffs = V.toList . unPoly

padTo :: Int -> [Fr] -> [Fr]
padTo n xs = take n (xs ++ repeat 0)

-- ------------------------------------------------------------------------------------------------------------------
-- verifyAssignment / verificationWitness[Zk] on a QAP value (the reference's types)
-- ------------------------------------------------------------------------------------------------------------------

-- | src/QAP.hs:300-327, one device call: a = d1*T + sum_k w_k L_k (b, c likewise), p = a*b - c, (h, rem) = p divMod T.
verificationWitnessZk :: Fr -> Fr -> Fr -> QAP Fr -> QapSet Fr -> Maybe (VPoly Fr)
verificationWitnessZk d1 d2 d3 QAP{..} assignment = unsafePerformIO $ do
  let lay   = layoutOf [void qapInputsLeft, void qapInputsRight, void qapOutputs, void assignment]
      sets  = [qapInputsLeft, qapInputsRight, qapOutputs]
      len   = maximum (1 : [ length (coeffs p) | s <- sets, p <- toList s ])
      flat s = frVector (concatMap (padTo len . coeffs) (dense 0 lay s))
      tgt   = coeffs qapTarget
      w     = frVector (dense 0 lay assignment)
  VS.unsafeWith (flat qapInputsLeft) $ \pl -> VS.unsafeWith (flat qapInputsRight) $ \pr ->
    VS.unsafeWith (flat qapOutputs) $ \po -> VS.unsafeWith (frVector tgt) $ \pt ->
    VS.unsafeWith w $ \pw -> VS.unsafeWith (frVector [d1, d2, d3]) $ \pd ->
    alloca $ \pq -> alloca $ \pn -> alloca $ \pdiv -> do
      c_qap_upload globalCtx pl pr po (fromIntegral (nCols lay)) (fromIntegral len) pt (fromIntegral (length tgt)) pq
        >>= ok "acg_qap_upload"      -- a zero target is ACG_ERR_BAD_ARG: the reference's quotRem divides by zero there
      q <- peek pq
      cap <- c_qap_quotient_len q
      r <- allocaBytes (32 * max 1 (fromIntegral cap)) $ \ph -> do
        c_qap_verify globalCtx q pw pd ph cap pn pdiv >>= ok "acg_qap_verify"
        divisible <- peek pdiv
        if divisible == 0 then pure Nothing else do
          n  <- peek pn
          ws <- peekArray (4 * fromIntegral n) ph
          pure . Just . toPoly . V.fromList . map unlimbs $ chunk 4 ws   -- toPoly strips trailing zeros
      c_qap_free q
      pure r

verificationWitness :: QAP Fr -> QapSet Fr -> Maybe (VPoly Fr)
verificationWitness = verificationWitnessZk 0 0 0

verifyAssignment :: QAP Fr -> QapSet Fr -> Bool
verifyAssignment qap assignment = isJust (verificationWitness qap assignment)

-- ------------------------------------------------------------------------------------------------------------------
-- building the QAP value: createPolynomials (Lagrange) and createPolynomialsFFT
-- ------------------------------------------------------------------------------------------------------------------

-- | src/QAP.hs:512-525: every wire's column (values in ascending-root order, zero padded to 2^k) is interpolated by
-- one batched inverse NTT on the GPU; the target is FFT.fftTargetPoly (acg_fft_target).  The library's transforms use
-- 'getRootOfUnity' of the field -- what every call site of the reference passes; for any other root function the
-- reference's own implementation is used.
createPolynomialsFFT :: (Int -> Fr) -> GenQAP (Map Fr) Fr -> QAP Fr
createPolynomialsFFT primRoots g@GenQAP{..}
  | primRoots logN /= getRootOfUnity logN = Ref.createPolynomialsFFT primRoots g
  | otherwise = unsafePerformIO $ do
      let sets  = [genQapInputsLeft, genQapInputsRight, genQapOutputs]
          cols  = concatMap (map (padTo bigN . Map.elems) . toList) sets
      buf <- VS.thaw (frVector (concat cols))          -- a private mutable copy: the transform is in place
      VSM.unsafeWith buf $ \p ->
        c_interpolate_columns globalCtx p (fromIntegral logN) (fromIntegral (length cols)) >>= ok "acg_interpolate_columns"
      out <- VS.toList <$> VS.unsafeFreeze buf
      tgt <- allocaBytes (32 * (n + 1)) $ \pt -> do
        c_fft_target globalCtx (fromIntegral n) pt >>= ok "acg_fft_target"
        map unlimbs . chunk 4 <$> peekArray (4 * (n + 1)) pt
      let polys = map (toPoly . V.fromList . map unlimbs . chunk 4) (chunk (4 * bigN) out)
          (l, ps1) = refillSet genQapInputsLeft polys
          (r, ps2) = refillSet genQapInputsRight ps1
          (o, _)   = refillSet genQapOutputs ps2
      pure QAP { qapInputsLeft = l, qapInputsRight = r, qapOutputs = o, qapTarget = toPoly (V.fromList tgt) }
  where
    n    = Map.size genQapTarget
    bigN = until (>= max 1 n) (* 2) 1
    logN = length (takeWhile (< bigN) (iterate (* 2) 1))

-- | src/QAP.hs:486-508: Lagrange interpolation through the GenQAP's own roots on the GPU (K5), target prod (X - root).
-- At most 4096 roots; beyond that the reference's implementation (O(m n^2)) is the only one there is.
createPolynomials :: GenQAP (Map Fr) Fr -> QAP Fr
createPolynomials g@GenQAP{..}
  | n == 0 || n > 4096 = Ref.createPolynomials g
  | otherwise = unsafePerformIO $ do
      let sets = [genQapInputsLeft, genQapInputsRight, genQapOutputs]
          xs   = Map.keys genQapTarget
          ys   = concatMap (map (\m -> [ Map.findWithDefault 0 x m | x <- xs ]) . toList) sets
          np   = length ys
      (cs, tgt) <- VS.unsafeWith (frVector xs) $ \px -> VS.unsafeWith (frVector (concat ys)) $ \py ->
        allocaBytes (32 * max 1 (np * n)) $ \pc -> allocaBytes (32 * (n + 1)) $ \pt -> do
          c_lagrange globalCtx px py (fromIntegral n) (fromIntegral np) pc pt >>= ok "acg_lagrange"
          (,) <$> peekArray (4 * np * n) pc <*> peekArray (4 * (n + 1)) pt
      let polys = map (toPoly . V.fromList . map unlimbs . chunk 4) (chunk (4 * n) cs)
          (l, ps1) = refillSet genQapInputsLeft polys
          (r, ps2) = refillSet genQapInputsRight ps1
          (o, _)   = refillSet genQapOutputs ps2
      pure QAP { qapInputsLeft = l, qapInputsRight = r, qapOutputs = o
               , qapTarget = toPoly (V.fromList (map unlimbs (chunk 4 tgt))) }
  where n = Map.size genQapTarget

arithCircuitToQAP :: [[Fr]] -> ArithCircuit Fr -> QAP Fr
arithCircuitToQAP roots circuit = createPolynomials (arithCircuitToGenQAP roots circuit)

arithCircuitToQAPFFT :: (Int -> Fr) -> [[Fr]] -> ArithCircuit Fr -> QAP Fr
arithCircuitToQAPFFT primRoots roots circuit = createPolynomialsFFT primRoots (arithCircuitToGenQAP roots circuit)

-- | src/QAP.hs:355-362 (needs createMapGenQap from the export-list patch).
gateToQAP :: (Int -> Fr) -> [Fr] -> Gate Wire Fr -> QAP Fr
gateToQAP primRoots roots = createPolynomialsFFT primRoots . addMissingZeroes roots . createMapGenQap . gateToGenQAP roots

-- ------------------------------------------------------------------------------------------------------------------
-- R1CS-form fast path (new names): the per-wire polynomials are never built
-- ------------------------------------------------------------------------------------------------------------------

-- | Transpose one QapSet of root->coefficient maps (column-major, src/QAP.hs:94-99) into CSR rows in
-- ascending-root order; explicit zeros are dropped.
toCsr :: Layout -> [Fr] -> QapSet (Map Fr Fr) -> (VS.Vector Word32, VS.Vector Word32, VS.Vector Word64)
toCsr lay roots s = (VS.fromList rowptr, VS.fromList cols, VS.fromList (concatMap limbs vals))
  where
    columns = zip [0 :: Int ..] (dense Map.empty lay s)
    rowOf r = [ (fromIntegral c, v) | (c, m) <- columns, Just v <- [Map.lookup r m], v /= 0 ]
    rows    = map rowOf roots
    rowptr  = scanl (+) 0 (map (fromIntegral . length) rows)
    (cols, vals) = unzip (concat rows)

-- | struct acg_csr { const uint32_t* rowptr; const uint32_t* col; const uint64_t* val; uint64_t nnz; }
withCsr :: (VS.Vector Word32, VS.Vector Word32, VS.Vector Word64) -> (Ptr () -> IO a) -> IO a
withCsr (rp, cl, vl) k =
  VS.unsafeWith rp $ \prp -> VS.unsafeWith cl $ \pcl -> VS.unsafeWith vl $ \pvl ->
  allocaBytes 32 $ \s -> do
    pokeByteOff s 0 prp
    pokeByteOff s 8 pcl
    pokeByteOff s 16 pvl
    pokeByteOff s 24 (fromIntegral (VS.length cl) :: Word64)
    k s

marshal :: GenQAP (Map Fr) Fr -> QapSet Fr
        -> (CUInt -> CUInt -> Ptr () -> Ptr () -> Ptr () -> Ptr Word64 -> IO a) -> IO a
marshal GenQAP{..} assignment k =
  withCsr (toCsr lay roots genQapInputsLeft) $ \a ->
  withCsr (toCsr lay roots genQapInputsRight) $ \b ->
  withCsr (toCsr lay roots genQapOutputs) $ \c ->
  VS.unsafeWith (frVector (dense 0 lay assignment)) $ \pw ->
    k (fromIntegral (length roots)) (fromIntegral (nCols lay)) a b c pw
  where
    roots = Map.keys genQapTarget
    lay   = layoutOf [void genQapInputsLeft, void genQapInputsRight, void genQapOutputs, void assignment]

-- | 'verifyAssignment' in R1CS form: valid iff every row satisfies (A.w)(B.w) = C.w
-- (equivalent to src/QAP.hs:276-282 because the target has distinct roots).
verifyAssignmentR1CS :: GenQAP (Map Fr) Fr -> QapSet Fr -> Bool
verifyAssignmentR1CS g assignment = unsafePerformIO $
  marshal g assignment $ \n m a b c pw -> alloca $ \pv -> alloca $ \pf -> do
    c_r1cs_check_host globalCtx n m a b c pw pv pf >>= ok "acg_r1cs_check_host"
    (== 0) <$> peek pv

verificationWitnessR1CS :: GenQAP (Map Fr) Fr -> QapSet Fr -> Maybe (VPoly Fr)
verificationWitnessR1CS = verificationWitnessZkR1CS 0 0 0

-- | src/QAP.hs:300-327 for the FFT-built QAP of this GenQAP (T = X^N - 1, N the next power of two of the root count),
-- by the linearity collapse sum_k w_k interpolate(col_k) = interpolate(A.w): 7 transforms instead of 3 (m + 1).
verificationWitnessZkR1CS :: Fr -> Fr -> Fr -> GenQAP (Map Fr) Fr -> QapSet Fr -> Maybe (VPoly Fr)
verificationWitnessZkR1CS d1 d2 d3 g assignment = unsafePerformIO $
  marshal g assignment $ \n m a b c pw ->
  alloca $ \pm -> alloca $ \pvec -> alloca $ \pdiv ->
  VS.unsafeWith (frVector [d1, d2, d3]) $ \pd -> do
    let bigN = until (>= fromIntegral n) (* 2) (1 :: Int)
    c_r1cs_upload globalCtx n m a b c 0 n pm >>= ok "acg_r1cs_upload"
    mh <- peek pm
    c_witness_upload globalCtx pw m pvec >>= ok "acg_witness_upload"
    vh <- peek pvec
    r <- allocaBytes (32 * (bigN + 1)) $ \ph -> do
      c_qap_witness globalCtx mh vh pd nullPtr nullPtr nullPtr ph pdiv >>= ok "acg_qap_witness"
      divisible <- peek pdiv
      if divisible == 0 then pure Nothing else do
        ws <- peekArray (4 * (bigN + 1)) ph
        pure . Just . toPoly . V.fromList . map unlimbs $ chunk 4 ws   -- toPoly strips trailing zeros
    c_vec_free vh
    c_r1cs_free mh
    pure r
