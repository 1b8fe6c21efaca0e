{-# LANGUAGE ForeignFunctionInterface, RecordWildCards, ScopedTypeVariables #-}

-- | Drop-in GPU path for the numeric core of "QAP" (sdiehl/arithmetic-circuits, src/QAP.hs).
--
-- This module is the reference-side binding of the C ABI in @include/acg.h@ (libacg.so, CUDA sm_100a).
-- It keeps the reference's names and argument order, monomorphised to BN254 'Fr' (the FFI cannot be
-- class-polymorphic; a second copy over BLS12-381 'Fr' differs only in 'fieldId'):
--
-- > generateAssignment   :: ArithCircuit Fr -> Map Int Fr -> QapSet Fr          -- unchanged (re-export)
-- > arithCircuitToGenQAP :: [[Fr]] -> ArithCircuit Fr -> GenQAP (Map Fr) Fr     -- unchanged (re-export)
-- > verifyAssignmentR1CS :: GenQAP (Map Fr) Fr -> QapSet Fr -> Bool             -- GPU: acg_r1cs_check_host
-- > verificationWitnessZk:: Fr -> Fr -> Fr -> GenQAP (Map Fr) Fr -> QapSet Fr -> Maybe (VPoly Fr)
-- >                                                                             -- GPU: acg_qap_witness
-- > createPolynomialsFFT :: (Int -> Fr) -> GenQAP (Map Fr) Fr -> QAP Fr          -- GPU: acg_interpolate_columns
--
-- NOTE: the build image of this repository has no GHC (SURVEY.md F3), so this file is shipped as source
-- and has not been compiled here; the same ABI is exercised from Python (ctypes) and C++ by the test
-- suite.  Marshalling follows include/acg.h: a field element is 4 little-endian Word64 limbs of @fromP@.
module QAP.GPU
  ( withAcg
  , verifyAssignmentR1CS
  , verifyAssignment
  , verificationWitness
  , verificationWitnessZk
  , createPolynomialsFFT
  , module QAP
  ) where

import Protolude hiding (quot, quotRem)

import           Data.Bits             (shiftL, shiftR, (.&.))
import qualified Data.Map              as Map
import           Data.Pairing.BN254    (Fr)
import           Data.Field.Galois     (fromP)
import           Data.Poly             (VPoly, toPoly)
import qualified Data.Vector           as V
import qualified Data.Vector.Storable  as VS
import           Foreign.C.Types       (CInt(..), CUInt(..), CULong(..))
import           Foreign.Marshal.Alloc (alloca, allocaBytes)
import           Foreign.Marshal.Array (peekArray)
import           Foreign.Ptr           (Ptr, nullPtr, castPtr)
import           Foreign.Storable      (peek, poke, pokeByteOff)
import           System.IO.Unsafe      (unsafePerformIO)

import QAP hiding (verifyAssignment, verificationWitness, verificationWitnessZk, createPolynomialsFFT)

-- opaque handles
data AcgCtx
data AcgR1cs
data AcgVec

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
-- scale-and-sum over a per-wire QAP (foldQapSet / combineWithDefaults, src/QAP.hs:163-181, 314-324):
-- polys (n_polys * len * 4 limbs), weights (n_polys * 4 limbs) -> out (len * 4 limbs)
foreign import ccall safe "acg_poly_combine"    c_poly_combine
  :: Ptr AcgCtx -> Ptr Word64 -> Ptr Word64 -> CUInt -> CUInt -> Ptr Word64 -> IO CInt
-- witness generation on the device (K6): circuit handle from acg_circuit_parse (the word stream a
-- `ArithCircuit Fr -> [Word64]` marshaller emits, include/acg.h), inputs as (index, limbs) pairs
data AcgCircuit
foreign import ccall safe "acg_circuit_parse"   c_circuit_parse   :: CInt -> Ptr Word64 -> Word64 -> Ptr (Ptr AcgCircuit) -> IO CInt
foreign import ccall safe "acg_circuit_free"    c_circuit_free    :: Ptr AcgCircuit -> IO ()
foreign import ccall safe "acg_generate_assignment_device" c_generate_assignment_device
  :: Ptr AcgCtx -> Ptr AcgCircuit -> Ptr Word32 -> Ptr Word64 -> CUInt -> CUInt -> CUInt -> CUInt
  -> Ptr (Ptr AcgVec) -> Ptr CUInt -> IO CInt
foreign import ccall safe "acg_vec_download"    c_vec_download    :: Ptr AcgCtx -> Ptr AcgVec -> Ptr Word64 -> CUInt -> IO CInt

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

-- | @fromP@ as 4 little-endian limbs.
limbs :: Fr -> [Word64]
limbs x = [ fromIntegral ((n `shiftR` (64 * i)) .&. 0xFFFFFFFFFFFFFFFF) | i <- [0 .. 3] ]
  where n = toInteger (fromP x)

unlimbs :: [Word64] -> Fr
unlimbs ws = fromInteger $ sum [ toInteger w `shiftL` (64 * i) | (w, i) <- zip ws [0 ..] ]

-- | Witness index layout of qapSetToMap (src/QAP.hs:605-620) with the block sizes of the GenQAP's wires.
data Layout = Layout { nIn, nMid, nOut :: Int }

layoutOf :: [QapSet a] -> Layout
layoutOf qs = Layout (mx qapSetInput) (mx qapSetIntermediate) (mx qapSetOutput)
  where mx f = maximum (0 : [ k + 1 | q <- qs, k <- Map.keys (f q) ])

-- | Transpose one QapSet of root->coefficient maps (column-major, src/QAP.hs:94-99) into CSR rows in
-- ascending-root order; explicit zeros are dropped.
toCsr :: Layout -> [Fr] -> QapSet (Map Fr Fr) -> (VS.Vector Word32, VS.Vector Word32, VS.Vector Word64)
toCsr Layout{..} roots QapSet{..} = (VS.fromList rowptr, VS.fromList cols, VS.fromList (concatMap limbs vals))
  where
    columns = (0, qapSetConstant)
            : [ (1 + k, m) | (k, m) <- Map.toList qapSetInput ]
           ++ [ (1 + nIn + k, m) | (k, m) <- Map.toList qapSetIntermediate ]
           ++ [ (1 + nIn + nMid + k, m) | (k, m) <- Map.toList qapSetOutput ]
    rowOf r = [ (fromIntegral c, v) | (c, m) <- columns, Just v <- [Map.lookup r m], v /= 0 ]
    rows    = map rowOf roots
    rowptr  = scanl (+) 0 (map (fromIntegral . length) rows)
    (cols, vals) = unzip (concat rows)

-- | Dense w in qapSetToMap order; wires the assignment lacks are 0 (src/QAP.hs:314 default).
witnessVector :: Layout -> QapSet Fr -> VS.Vector Word64
witnessVector Layout{..} QapSet{..} = VS.fromList (concatMap limbs dense)
  where
    n = 1 + nIn + nMid + nOut
    m = Map.fromList $ (0, qapSetConstant)
          : [ (1 + k, v) | (k, v) <- Map.toList qapSetInput ]
         ++ [ (1 + nIn + k, v) | (k, v) <- Map.toList qapSetIntermediate ]
         ++ [ (1 + nIn + nMid + k, v) | (k, v) <- Map.toList qapSetOutput ]
    dense = [ Map.findWithDefault 0 i m | i <- [0 .. n - 1] ]

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
  VS.unsafeWith w $ \pw ->
    k (fromIntegral (length roots)) (fromIntegral (1 + nIn lay + nMid lay + nOut lay)) a b c pw
  where
    roots = Map.keys genQapTarget
    lay   = layoutOf [void genQapInputsLeft, void genQapInputsRight, void genQapOutputs, void assignment]
    w     = witnessVector lay assignment

-- | 'verifyAssignment' in R1CS form: valid iff every row satisfies (A.w)(B.w) = C.w
-- (equivalent to src/QAP.hs:276-282 because the target has distinct roots).
verifyAssignmentR1CS :: GenQAP (Map Fr) Fr -> QapSet Fr -> Bool
verifyAssignmentR1CS g assignment = unsafePerformIO $
  marshal g assignment $ \n m a b c pw -> alloca $ \pv -> alloca $ \pf -> do
    rc <- c_r1cs_check_host globalCtx n m a b c pw pv pf
    when (rc /= 0) $ panic ("acg_r1cs_check_host failed: " <> show rc)
    (== 0) <$> peek pv

-- | Same name and meaning as the reference; takes the GenQAP (the polynomial QAP is never needed).
verifyAssignment :: GenQAP (Map Fr) Fr -> QapSet Fr -> Bool
verifyAssignment = verifyAssignmentR1CS

verificationWitness :: GenQAP (Map Fr) Fr -> QapSet Fr -> Maybe (VPoly Fr)
verificationWitness = verificationWitnessZk 0 0 0

-- | src/QAP.hs:300-327 on the FFT-built QAP (T = X^N - 1, N the next power of two of the root count).
verificationWitnessZk :: Fr -> Fr -> Fr -> GenQAP (Map Fr) Fr -> QapSet Fr -> Maybe (VPoly Fr)
verificationWitnessZk d1 d2 d3 g assignment = unsafePerformIO $
  marshal g assignment $ \n m a b c pw ->
  alloca $ \pm -> alloca $ \pvec -> alloca $ \pdiv ->
  VS.unsafeWith (VS.fromList (concatMap limbs [d1, d2, d3])) $ \pd -> do
    let bigN = until (>= fromIntegral n) (* 2) (1 :: Int)
    rc1 <- c_r1cs_upload globalCtx n m a b c 0 n pm
    when (rc1 /= 0) $ panic ("acg_r1cs_upload failed: " <> show rc1)
    mh <- peek pm
    rc2 <- c_witness_upload globalCtx pw m pvec
    when (rc2 /= 0) $ panic ("acg_witness_upload failed: " <> show rc2)
    vh <- peek pvec
    r <- allocaBytes (32 * (bigN + 1)) $ \ph -> do
      rc <- c_qap_witness globalCtx mh vh pd nullPtr nullPtr nullPtr ph pdiv
      when (rc /= 0) $ panic ("acg_qap_witness failed: " <> show rc)
      ok <- peek pdiv
      if ok == 0 then pure Nothing else do
        ws <- peekArray (4 * (bigN + 1)) ph
        pure . Just . toPoly . V.fromList . map unlimbs $ chunks4 ws   -- toPoly strips trailing zeros
    c_vec_free vh
    c_r1cs_free mh
    pure r
  where
    chunks4 [] = []
    chunks4 xs = let (h, t) = splitAt 4 xs in h : chunks4 t

-- | src/QAP.hs:512-525: every wire's column (values in ascending-root order, zero padded to 2^k) is
-- interpolated by one batched inverse NTT on the GPU.  The primitive-root function is ignored: the
-- library uses getRootOfUnity of the field, which is what every call site passes.
createPolynomialsFFT :: (Int -> Fr) -> GenQAP (Map Fr) Fr -> QAP Fr
createPolynomialsFFT _primRoots GenQAP{..} = unsafePerformIO $ do
  let bigN   = until (>= Map.size genQapTarget) (* 2) 1
      logN   = length (takeWhile (< bigN) (iterate (* 2) 1))
      pad xs = take bigN (xs ++ repeat 0)
      sets   = [genQapInputsLeft, genQapInputsRight, genQapOutputs]
      cols   = concatMap (map (pad . Map.elems) . toList) sets
      flat   = VS.fromList (concatMap limbs (concat cols))
  out <- VS.unsafeWith flat $ \p -> do   -- in place on a private copy
    rc <- c_interpolate_columns globalCtx (castPtr p) (fromIntegral logN) (fromIntegral (length cols))
    when (rc /= 0) $ panic ("acg_interpolate_columns failed: " <> show rc)
    peekArray (VS.length flat) p
  let polys     = map (toPoly . V.fromList . map unlimbs . chunk 4) (chunk (4 * bigN) out)
      refill q ps = let (here, rest) = splitAt (length (toList q)) ps in (fill q here, rest)
      fill q ps = snd (mapAccumL (\(x : xs) _ -> (xs, x)) ps q)
      (l, ps1)  = refill genQapInputsLeft polys
      (r, ps2)  = refill genQapInputsRight ps1
      (o, _)    = refill genQapOutputs ps2
  pure QAP { qapInputsLeft = l, qapInputsRight = r, qapOutputs = o
           , qapTarget = toPoly (V.fromList (negate 1 : replicate (bigN - 1) 0 ++ [1])) }  -- X^N - 1
  where
    chunk _ [] = []
    chunk k xs = let (h, t) = splitAt k xs in h : chunk k t
