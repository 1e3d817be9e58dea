! implsch_b200.F90 -- IMPLSCH with the reference's call signature (src/ecwam/implsch.F90:10-23, dummies as :117-143), body on
! the B200: replaces src/ecwam/implsch.F90 in an ecWAM build that links libecwam_b200.so.
!
! WAMINTGR calls it per NPROMA chunk with the chunk's slices of the FIELD_API arrays (wamintgr.F90:117-146).  The actual
! arguments must be the DEVICE copies (wamintgr_loki_gpu.F90:141-157: GET_DEVICE_DATA_*), which is what the GPU build of WAMINTGR
! passes; their addresses reach the library through !$acc host_data use_device, the idiom of mpexchng.F90:170-174.  The library
! checks that every argument is the same chunk of the arrays bound at set-up (fortran/ecwam_b200_setup.F90) and derives ICHNK
! from FL1's address.  Per-chunk launches are correct but small: the production path is the chunk loop as ONE call,
! ECWAM_B200_IMPLSCH_ALL / ECWAM_B200_WAMINTGR (see the note at the end of this file).

SUBROUTINE IMPLSCH (KIJS, KIJL, FL1,                         &
 &                  WAVNUM, CGROUP, CIWA, CINV, XK2CG, STOKFAC, &
 &                  EMAXDPT, DEPTH, IOBND, IODP,IBRMEM,      &
 &                  AIRD, WDWAVE, CICOVER, WSWAVE, WSTAR, USTRA, VSTRA, &
 &                  UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, CITHICK, &
 &                  NEMOUSTOKES, NEMOVSTOKES, NEMOSTRN, &
 &                  NPHIEPS, NTAUOC, NSWH, NMWP, NEMOTAUX, &
 &                  NEMOTAUY, NEMOTAUICX, NEMOTAUICY, &
 &                  NEMOWSWAVE, NEMOPHIF, &
 &                  WSEMEAN, WSFMEAN, USTOKES, VSTOKES, STRNMS, &
 &                  TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, &
 &                  TAUICX, TAUICY, &
 &                  PHIOCD, PHIEPS, PHIAW, &
 &                  MIJ, XLLWS)

      USE PARKIND_WAVE, ONLY : JWIM, JWRB, JWRO
      USE YOWPARAM , ONLY : NANG     ,NFRE
      USE YOWABORT , ONLY : WAM_ABORT
      USE YOMHOOK  , ONLY : LHOOK,   DR_HOOK, JPHOOK
      USE ECWAM_B200_MOD, ONLY : B200_HANDLE, ECWAM_B200_IMPLSCH_F, ECWAM_B200_ERRMSG
      USE, INTRINSIC :: ISO_C_BINDING, ONLY : C_INT

      IMPLICIT NONE

      INTEGER(KIND=JWIM), INTENT(IN) :: KIJS, KIJL
      REAL(KIND=JWRB), DIMENSION(KIJL,NANG,NFRE), INTENT(INOUT) :: FL1
      REAL(KIND=JWRB), DIMENSION(KIJL, NFRE), INTENT(IN) :: WAVNUM
      REAL(KIND=JWRB), DIMENSION(KIJL, NFRE), INTENT(IN) :: CGROUP
      REAL(KIND=JWRB), DIMENSION(KIJL, NFRE), INTENT(IN) :: CIWA
      REAL(KIND=JWRB), DIMENSION(KIJL, NFRE), INTENT(IN) :: CINV
      REAL(KIND=JWRB), DIMENSION(KIJL, NFRE), INTENT(IN) :: XK2CG
      REAL(KIND=JWRB), DIMENSION(KIJL, NFRE), INTENT(IN) :: STOKFAC

      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(IN) :: EMAXDPT
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(IN) :: DEPTH
      INTEGER(KIND=JWIM), DIMENSION(KIJL), INTENT(IN) :: IODP
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(IN) :: IBRMEM
      INTEGER(KIND=JWIM), DIMENSION(KIJL), INTENT(IN) :: IOBND

      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(INOUT) :: AIRD, WDWAVE, CICOVER, WSWAVE, WSTAR, USTRA, VSTRA
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(INOUT) :: UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, CITHICK
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(INOUT) :: WSEMEAN, WSFMEAN, USTOKES, VSTOKES, STRNMS
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(INOUT) :: TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, PHIOCD
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(INOUT) :: TAUICX, TAUICY
      REAL(KIND=JWRB), DIMENSION(KIJL), INTENT(INOUT) :: PHIEPS, PHIAW
      REAL(KIND=JWRO), DIMENSION(KIJL), INTENT(INOUT) :: NEMOUSTOKES, NEMOVSTOKES, NEMOSTRN
      REAL(KIND=JWRO), DIMENSION(KIJL), INTENT(INOUT) :: NPHIEPS, NTAUOC, NSWH, NMWP, NEMOTAUX
      REAL(KIND=JWRO), DIMENSION(KIJL), INTENT(INOUT) :: NEMOTAUY, NEMOWSWAVE, NEMOPHIF
      REAL(KIND=JWRO), DIMENSION(KIJL), INTENT(INOUT) :: NEMOTAUICX, NEMOTAUICY
      INTEGER(KIND=JWIM), DIMENSION(KIJL), INTENT(OUT) :: MIJ
      REAL(KIND=JWRB), DIMENSION(KIJL,NANG,NFRE), INTENT(OUT) :: XLLWS

      INTEGER(KIND=C_INT) :: IERR
      REAL(KIND=JPHOOK) :: ZHOOK_HANDLE

! ----------------------------------------------------------------------

      IF (LHOOK) CALL DR_HOOK('IMPLSCH',0,ZHOOK_HANDLE)

!$acc host_data use_device(FL1, WAVNUM, CGROUP, CIWA, CINV, XK2CG, STOKFAC, EMAXDPT, DEPTH, IOBND, IODP, IBRMEM, &
!$acc &   AIRD, WDWAVE, CICOVER, WSWAVE, WSTAR, USTRA, VSTRA, UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, CITHICK, &
!$acc &   NEMOUSTOKES, NEMOVSTOKES, NEMOSTRN, NPHIEPS, NTAUOC, NSWH, NMWP, NEMOTAUX, NEMOTAUY, NEMOTAUICX, NEMOTAUICY, &
!$acc &   NEMOWSWAVE, NEMOPHIF, WSEMEAN, WSFMEAN, USTOKES, VSTOKES, STRNMS, TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, &
!$acc &   TAUICX, TAUICY, PHIOCD, PHIEPS, PHIAW, MIJ, XLLWS)
      IERR = ECWAM_B200_IMPLSCH_F(B200_HANDLE, INT(KIJS,C_INT), INT(KIJL,C_INT), FL1, &
 &                  WAVNUM, CGROUP, CIWA, CINV, XK2CG, STOKFAC, &
 &                  EMAXDPT, DEPTH, IOBND, IODP, IBRMEM, &
 &                  AIRD, WDWAVE, CICOVER, WSWAVE, WSTAR, USTRA, VSTRA, &
 &                  UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, CITHICK, &
 &                  NEMOUSTOKES, NEMOVSTOKES, NEMOSTRN, &
 &                  NPHIEPS, NTAUOC, NSWH, NMWP, NEMOTAUX, &
 &                  NEMOTAUY, NEMOTAUICX, NEMOTAUICY, &
 &                  NEMOWSWAVE, NEMOPHIF, &
 &                  WSEMEAN, WSFMEAN, USTOKES, VSTOKES, STRNMS, &
 &                  TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, &
 &                  TAUICX, TAUICY, &
 &                  PHIOCD, PHIEPS, PHIAW, &
 &                  MIJ, XLLWS)
!$acc end host_data

!     The reference aborts on errors (no return codes, e.g. sinput_ard.F90:286-295)
      IF (IERR /= 0) CALL WAM_ABORT('IMPLSCH (ecwam_b200): '//TRIM(ECWAM_B200_ERRMSG()),__FILENAME__,__LINE__)

      IF (LHOOK) CALL DR_HOOK('IMPLSCH',1,ZHOOK_HANDLE)

END SUBROUTINE IMPLSCH

! Note for WAMINTGR (wamintgr.F90:117-146 / wamintgr_loki_gpu.F90:164-189): the chunk loop
!     DO ICHNK=1,NCHNK
!       CALL IMPLSCH (1, NPROMA_WAM, VARS_4D%FL1(:,:,:,ICHNK), WVPRPT%WAVNUM(:,:,ICHNK), ... )
!     ENDDO
! is one library call that processes every chunk in one launch sequence,
!     IERR = ECWAM_B200_IMPLSCH_ALL(B200_HANDLE)
! and, when the propagation is due in the same step (CDATE == CDTPRA, IDELPRO == IDELT), PROPAG_WAM + the loop are
!     IERR = ECWAM_B200_WAMINTGR(B200_HANDLE)
! (same results; the block -> chunk copy of PROPAG_WAM is folded into IMPLSCH's loads).
